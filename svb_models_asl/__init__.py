"""
Drop-in alias: ``from svb_models_asl import AslRestModel, AslRestDisp, AslNNModel``
(/root/reference/svb_models_asl/__init__.py:7-16) resolves to the B200-native plugins.
"""
from svb_models_asl_b200 import __version__  # noqa: F401
from svb_models_asl_b200.plugin import *  # noqa: F401,F403
from svb_models_asl_b200.plugin import MODELS as _MODELS

globals().update({cls.__name__: cls for cls in _MODELS.values()})
__all__ = [cls.__name__ for cls in _MODELS.values()] + ["__version__"]
