"""Alias of svb_models_asl_b200.svbcompat.utils (see svb/__init__.py)."""
from svb_models_asl_b200.svbcompat.utils import *  # noqa: F401,F403
from svb_models_asl_b200.svbcompat import utils as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
