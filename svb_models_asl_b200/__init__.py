"""
svb_models_asl_b200 - B200-native kernels and host layer for the ASL stochastic-variational-Bayes hot path
of physimals/svb_models_asl.  The CUDA library is loaded lazily by the operators; importing
this package on a machine without a GPU works (descriptors, option handling, I/O) but any compute call raises.
"""
from .plugin import MODELS, get_model_class  # noqa: F401
from .plugin.aslrest import AslRestModel  # noqa: F401
from .plugin.aslnn import AslNNModel  # noqa: F401
from .plugin.aslrest_disp import AslRestDisp  # noqa: F401

__version__ = "0.1.0+b200"
__all__ = ["AslRestModel", "AslRestDisp", "AslNNModel", "MODELS", "get_model_class", "__version__"]
