"""Alias of svb_models_asl_b200.svbcompat.dist (see svb/__init__.py)."""
from svb_models_asl_b200.svbcompat.dist import *  # noqa: F401,F403
from svb_models_asl_b200.svbcompat import dist as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
