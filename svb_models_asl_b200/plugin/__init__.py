"""The three `svb.models` entry points of the reference (setup.py:89-95), backed by libsvbasl.so."""
from .aslnn import AslNNModel  # noqa: F401
from .aslrest import AslRestModel  # noqa: F401
from .aslrest_disp import AslRestDisp  # noqa: F401

MODELS = {"aslrest": AslRestModel, "aslrest_disp": AslRestDisp, "aslnn": AslNNModel}


def get_model_class(name):
    """What svb does through the `svb.models` entry-point group (setup.py:89-95)."""
    if name not in MODELS:
        raise ValueError("No such model: %s (known: %s)" % (name, ", ".join(sorted(MODELS))))
    return MODELS[name]
