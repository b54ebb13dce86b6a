"""
``aslrest_disp`` plugin: resting-state ASL model with gamma-kernel bolus dispersion, evaluated as the numerical
convolution of the dispersed AIF with the well-mixed residue function.

Host-side mirror of ``AslRestDisp`` (/root/reference/svb_models_asl/aslrest_disp.py): the parent's options plus
``conv_dt``, ``conv_type``, ``infer_disp_params`` (:24-28), the two LogNormal dispersion parameters ``s`` (7.4, var 2)
and ``sp`` (0.74, var 2) appended after the parent's parameters (:32-38), and the convolution grid (:41-43).  The
arithmetic - incomplete gamma functions, the convolution (as an O(NT) recurrence), the interpolation at the time
points and all derivatives - is csrc/model_disp.h.

The file as shipped cannot run; decisions: the parent's call signature is honoured,
the residue is per voxel, ``pvgm`` is applied, and the post-bolus AIF is the intended ``kc*(gamma2 - gamma1)``.
``disp_postbolus="as_written"`` reproduces the shipped ``gamma2 - gamma2 == 0`` (:108).
"""
import numpy as np

from ..svbcompat.model import ModelOption
from ..svbcompat.parameter import get_parameter
from .. import _lib as L
from .aslrest import AslRestModel, __version__


class AslRestDisp(AslRestModel):
    """ASL resting state model with explicit AIF (x) residue convolution, to incorporate dispersion"""

    OPTIONS = AslRestModel.OPTIONS + [
        ModelOption("conv_dt", "Time interval for numerical convolution", units="s", type=float, default=0.1),
        ModelOption("conv_type", "Convolution type ('gamma' only supprted type presently)", type=str, default="gamma"),
        ModelOption("infer_disp_params", "Whether to infer parameters of the dispersion", type=bool, default=True),
        ModelOption("disp_postbolus", "Post-bolus AIF: 'intended' (gamma2-gamma1) or 'as_written' (zero)", type=str,
                    default="intended"),
    ]

    KIND = L.MODEL_ASLREST_DISP

    def __init__(self, data_model, **options):
        AslRestModel.__init__(self, data_model, **options)
        if self.conv_type != "gamma":
            raise ValueError("Only gamma dispersion is supported (conv_type=%r)" % (self.conv_type,))
        if self.incwm or self.infert1:
            raise NotImplementedError("aslrest_disp: WM component / T1 inference are not available with dispersion "
                                      "(the reference raises on extra tissue parameters, aslrest.py:347-348)")
        if self.infer_disp_params:
            self.params.append(get_parameter("s", dist="LogNormal", mean=7.4, var=2.0, **options))
            self.params.append(get_parameter("sp", dist="LogNormal", mean=0.74, var=2.0, **options))
        # grid for the numerical convolution (aslrest_disp.py:41-43)
        self.conv_tmax = max(max(self.tis), 5.0)
        self.conv_nt = 1 + int(self.conv_tmax / self.conv_dt)
        self.conv_t = np.linspace(0.0, self.conv_tmax, self.conv_nt)

    def __str__(self):
        return "ASL resting state model with gamma dispersion: %s" % __version__

    def kernel_flags(self):
        f = AslRestModel.kernel_flags(self)
        if self.infer_disp_params:
            f |= L.F_DISP_INFER
        if self.disp_postbolus == "as_written":
            f |= L.F_DISP_ASWRITTEN
        return f

    def kernel_model(self, device_array=None):
        m, keep = AslRestModel.kernel_model(self, device_array)
        m.conv_dt, m.conv_tmax, m.conv_nt = self.conv_dt, self.conv_tmax, self.conv_nt
        m.s_fixed, m.sp_fixed = 7.4, 0.74                                # prior defaults from Fabber (:88-90)
        return m, keep
