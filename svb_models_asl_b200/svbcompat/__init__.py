"""
svb-compatible host layer for the ASL hot path.

Exposes the names the reference plugins and scripts import from the external
``svb`` package: ``DataModel``, ``__version__``, and the submodules
``model, utils, parameter, dist, prior, main``.
"""
__version__ = "0.1.0+b200"

from .data import DataModel  # noqa: E402

VolumetricModel = DataModel  # aslnn.py:22-25 accepts either name

__all__ = ["DataModel", "VolumetricModel", "__version__"]
