"""``from svb.main import run`` (scripts/asl_example.py:16)."""
from svb_models_asl_b200.svbcompat.main import run, main  # noqa: F401
