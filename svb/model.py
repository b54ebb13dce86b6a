"""Alias of svb_models_asl_b200.svbcompat.model (see svb/__init__.py)."""
from svb_models_asl_b200.svbcompat.model import *  # noqa: F401,F403
from svb_models_asl_b200.svbcompat import model as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
