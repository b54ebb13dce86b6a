"""
Drop-in alias for the external ``svb`` package names the reference imports
(``/root/reference/scripts/asl_example.py:16``, ``gen_test_data.py:10``,
``svb_models_asl/aslrest.py:11-13``).  Implementation: ``svb_models_asl_b200.svbcompat``.
"""
from svb_models_asl_b200.svbcompat import DataModel, VolumetricModel, __version__  # noqa: F401
