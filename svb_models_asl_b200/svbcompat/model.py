"""
Plugin base class: mirror of ``svb.model.Model`` / ``ModelOption`` which the
reference subclasses (``/root/reference/svb_models_asl/aslrest.py:11,17,24-70``).

Contract kept: ``OPTIONS`` is a list of ``ModelOption``; ``__init__`` copies each
option from the keyword arguments (or its default) onto ``self`` and ignores
unknown keys; ``params`` is the ordered parameter list; ``tpts()`` returns the
time points; ``evaluate(params, tpts)`` is the forward model; ``ievaluate``
returns numpy.
"""
import numpy as np

from .utils import LogBase, ValueList  # noqa: F401  (aslnn.py:20 imports ValueList from here)


class ModelOption:
    def __init__(self, attr_name, desc, **kwargs):
        self.attr_name = attr_name
        self.desc = desc
        self.clargs = kwargs.get("clargs", ["--%s" % attr_name.replace("_", "-")])
        self.default = kwargs.get("default", None)
        self.units = kwargs.get("units", None)
        self.type = kwargs.get("type", str)


class Model(LogBase):
    OPTIONS = [
        ModelOption("dt", "Time separation between volumes", type=float, default=1.0),
        ModelOption("t0", "Time offset for first volume", type=float, default=0.0),
    ]

    def __init__(self, data_model, **options):
        LogBase.__init__(self)
        self.data_model = data_model
        self.params = []
        for option in self.OPTIONS:
            setattr(self, option.attr_name, options.get(option.attr_name, option.default))

    @property
    def nparams(self):
        return len(self.params)

    def param_idx(self, name):
        for idx, param in enumerate(self.params):
            if param.name == name:
                return idx
        raise ValueError("Parameter not found in model: %s" % name)

    def tpts(self):
        n = self.data_model.n_tpts
        return np.linspace(self.t0, self.t0 + self.dt * n, num=n, endpoint=False, dtype=np.float32)

    def evaluate(self, params, tpts):
        raise NotImplementedError("evaluate")

    def ievaluate(self, params, tpts):
        """One-shot forward evaluation returning numpy (gen_test_data.py:47)."""
        out = self.evaluate(params, tpts)
        if hasattr(out, "detach"):
            out = out.detach().cpu().numpy()
        return np.asarray(out)

    def __str__(self):
        return "%s" % type(self).__name__
