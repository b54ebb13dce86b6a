"""The three `svb.models` entry points of the reference (setup.py:89-95), backed by libsvbasl.so."""
from .aslnn import AslNNModel  # noqa: F401
from .aslrest import AslRestModel  # noqa: F401
from .aslrest_disp import AslRestDisp  # noqa: F401

MODELS = {"aslrest": AslRestModel, "aslrest_disp": AslRestDisp, "aslnn": AslNNModel}


def get_model_class(name):
    """What svb does through the `svb.models` entry-point group (setup.py:89-95): installed distributions that
    register models under that group are found by name (pyproject.toml registers these three the same way); the
    built-in table answers first, so an uninstalled checkout works too."""
    if name in MODELS:
        return MODELS[name]
    try:
        from importlib.metadata import entry_points
        for ep in entry_points(group="svb.models"):
            if ep.name == name:
                return ep.load()
    except Exception:  # noqa: BLE001 - discovery is best effort; the error below names what is known
        pass
    raise ValueError("No such model: %s (known: %s)" % (name, ", ".join(sorted(MODELS))))
