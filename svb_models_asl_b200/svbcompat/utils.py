"""``svb.utils`` names the reference imports (``aslrest.py:12``)."""
import logging

import numpy as np

NP_DTYPE = np.float32


class ValueList:
    """Option type: comma/space separated string (or sequence) -> list of values."""
    def __init__(self, value_type=str):
        self._type = value_type

    def __call__(self, value):
        if isinstance(value, str):
            value = value.replace(",", " ").split()
        elif not isinstance(value, (list, tuple, np.ndarray)):
            value = [value]
        return [self._type(v) for v in value]


class LogBase:
    """Gives every object a ``self.log`` named after its class."""
    def __init__(self, **_kw):
        self.log = logging.getLogger(type(self).__name__)

    def log_tf(self, value, *_a, **_kw):
        # The reference wraps intermediates in this debug pass-through
        # (aslrest.py:269-340); there is no graph here so it is the identity.
        return value
