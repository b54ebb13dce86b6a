"""
``aslrest`` plugin: resting-state ASL kinetic model (Buxton; PASL / pCASL; tissue, optional WM
partial-volume component, optional arterial component; optional T1 inference).

Host-side mirror of ``AslRestModel`` (/root/reference/svb_models_asl/aslrest.py): same option names
and defaults (:24-67), same option resolution (:69-139), same parameter set, order, priors and
initialisers (:183-246, :461-520), same ``tpts()`` (:432-456) and ``evaluate(params, tpts)`` contract
(:248-340) - but the arithmetic runs in the CUDA kernels of libsvbasl.so (csrc/model_aslrest.h); there
is no TensorFlow graph and no CPU path.  ``kernel_model()`` gives the flat descriptor the fused
ELBO/gradient kernel consumes.

Scope notes: volumetric data only (surface/hybrid node spaces need the toblerone projection
of an svb fork that is not part of the reference tree); t1/pc/fcalib/att are uniform (the reference only
makes them per-node arrays for the surface mode), partial volumes may be per voxel.
"""
import numpy as np

from ..svbcompat.model import Model, ModelOption
from ..svbcompat.parameter import get_parameter
from ..svbcompat.utils import NP_DTYPE, ValueList
from .. import _lib as L

__version__ = "0.1.0+b200"


def _opt(name, desc, **kw):
    return ModelOption(name, desc, **kw)


class AslRestModel(Model):
    """ASL resting state model"""

    OPTIONS = [
        # acquisition
        _opt("tau", "Bolus duration", units="s", clargs=("--tau", "--bolus"), type=float, default=1.8),
        _opt("casl", "Data is CASL/pCASL", type=bool, default=False),
        _opt("tis", "Inversion times", units="s", type=ValueList(float)),
        _opt("plds", "Post-labelling delays (for CASL instead of TIs)", units="s", type=ValueList(float)),
        _opt("repeats", "Number of repeats - single value or one per TI/PLD", units="s", type=ValueList(int),
             default=[1]),
        _opt("slicedt", "Increase in TI/PLD per slice", units="s", type=float, default=0),
        # grey matter
        _opt("t1", "Tissue T1 value", units="s", type=float, default=1.3),
        _opt("att", "Bolus arrival time", units="s", clargs=("--bat",), type=float, default=1.3),
        _opt("attsd", "Bolus arrival time prior std.dev.", units="s", clargs=("--batsd",), type=float, default=None),
        _opt("fcalib", "Perfusion value to use in estimation of effective T1", type=float, default=0.01),
        _opt("pc", "Blood/tissue partition coefficient (default 0.9, or 0.98 when WM is included)", type=float,
             default=None),
        # white matter
        _opt("incwm", "Include WM parameters", default=False),
        _opt("fwm", "WM perfusion", type=float, default=0),
        _opt("attwm", "WM arterial transit time", clargs=("--batwm",), type=float, default=1.6),
        _opt("t1wm", "WM T1 value", units="s", type=float, default=1.1),
        _opt("pcwm", "WM parition coefficient", type=float, default=0.8),
        _opt("fcalibwm", "WM perfusion value to use in estimation of effective T1", type=float, default=0.003),
        # blood
        _opt("t1b", "Blood T1 value", units="s", type=float, default=1.65),
        _opt("artt", "Arterial bolus arrival time", units="s", clargs=("--batart",), type=float, default=None),
        _opt("arttsd", "Arterial bolus arrival time prior std.dev.", units="s", clargs=("--batartsd",), type=float,
             default=None),
        # what to infer
        _opt("inferatt", "Infer ATT (default on for multi-time imaging)", type=bool, default=None),
        _opt("artonly", "Only infer arterial component not tissue", type=bool),
        _opt("inferart", "Infer arterial component", type=bool),
        _opt("infert1", "Infer T1 value", type=bool),
        _opt("att_init", "Initialization method for ATT (max=max signal - bolus duration)", default=""),
        _opt("pvcorr", "Perform PVEc (shortcut for incwm, inferwm)", default=False),
        _opt("inferwm", "Infer WM parameters", default=False),
        # partial volumes
        _opt("pvgm", "GM partial volume", type=float, default=1.0),
        _opt("pvwm", "WM partial volume", type=float, default=0.0),
    ]

    KIND = L.MODEL_ASLREST

    def __init__(self, data_model, **options):
        Model.__init__(self, data_model, **options)
        self._resolve_timing()
        self._resolve_inference_flags()
        self._resolve_partial_volumes()
        self.leadscale = 0.01                                            # aslrest.py:232
        self._build_params(options)

    # ---- option resolution (aslrest.py:71-139) ----
    def _resolve_timing(self):
        if self.plds is not None:
            self.tis = [self.tau + pld for pld in self.plds]
        if self.tis is None:
            raise ValueError("Either TIs or PLDs must be given")
        self.tis = [float(t) for t in self.tis]
        multi = len(self.tis) > 1
        if self.inferatt is None:
            self.inferatt = multi
        elif not isinstance(self.inferatt, bool):
            raise ValueError("inferatt argument must be bool")
        if self.attsd is None:
            self.attsd = 1.0 if multi else 0.1
        if self.artt is None:
            self.artt = self.att - 0.3
        if self.arttsd is None:
            self.arttsd = self.attsd
        reps = self.repeats
        if isinstance(reps, (int, np.integer)):
            reps = [int(reps)]
        reps = list(reps)
        if len(reps) > 1 and any(r != reps[0] for r in reps):
            raise NotImplementedError("Variable repeats for TIs/PLDs")
        self.repeats = int(reps[0])

    def _resolve_inference_flags(self):
        if self.pvcorr:
            self.incwm = self.inferwm = True
        self.incwm = bool(self.incwm)        # inferwm alone adds the WM parameters but no WM signal (aslrest.py:327)
        if self.artonly:
            self.inferart = True
        self.inferart, self.infert1 = bool(self.inferart), bool(self.infert1)
        self.artonly, self.inferwm = bool(self.artonly), bool(self.inferwm)
        if not self.data_model.is_volumetric:
            raise NotImplementedError("surface / hybrid node spaces are outside the scope of this engine")
        if self.pc is None:
            self.pc = 0.98 if self.incwm else 0.9                        # aslrest.py:131-135

    def _resolve_partial_volumes(self):
        """pvgm/pvwm: scalar, per-voxel array or image file (aslrest.py:103-120)."""
        for name in ("pvgm", "pvwm"):
            value = getattr(self, name)
            if not isinstance(value, (int, float)):
                try:
                    arr = np.asarray(self.data_model._get_data(value)[1], dtype=NP_DTYPE).reshape(-1)
                except Exception as exc:  # noqa: BLE001
                    raise ValueError("Could not interpret PV estimates") from exc
                if arr.size == self.data_model.mask_flattened.size and arr.size != self.data_model.n_nodes:
                    arr = arr[self.data_model.mask_flattened]
                if arr.size == 1:
                    arr = float(arr[0])
                elif arr.size != self.data_model.n_nodes:
                    raise ValueError("Could not interpret PV estimates")
                setattr(self, name, arr)
        if self.incwm and (np.asarray(self.pvgm) + np.asarray(self.pvwm) > 1).any():
            raise ValueError("At least one GM and WM PV sum to > 1")

    # ---- parameters: order matters (aslrest.py:181-246) ----
    def _build_params(self, options):
        att_var = self.attsd ** 2
        flow = dict(dist="Normal", prior_var=1e6, post_var=1.5, post_init=self._init_flow)
        self.params = []
        if not self.artonly:
            self.params.append(get_parameter("ftiss", mean=1.5, **flow, **options))
            if self.inferatt:
                self.params.append(get_parameter("delttiss", dist="Normal", mean=self.att, var=att_var,
                                                 post_init=self._init_delt, **options))
            if self.inferwm:
                self.params.append(get_parameter("fwm", mean=0.5, **flow, **options))
                if self.inferatt:
                    self.params.append(get_parameter("deltwm", dist="Normal", mean=self.attwm, var=att_var,
                                                     post_init=self._init_delt, **options))
        if self.infert1:
            self.params.append(get_parameter("t1", mean=self.t1, var=0.01, **options))
            if self.inferwm:
                self.params.append(get_parameter("t1wm", mean=self.t1wm, var=0.01, **options))
        if self.inferart:
            self.params.append(get_parameter("fblood", dist="Normal", mean=0.0, prior_var=1e6, post_var=1.5,
                                             post_init=self._init_fblood, prior_type="A", **options))
            if self.inferatt:
                self.params.append(get_parameter("deltblood", dist="Normal", mean=self.artt, var=self.arttsd ** 2,
                                                 post_init=self._init_delt, **options))

    # ---- kernel descriptor ----
    def kernel_flags(self):
        f = 0
        for on, bit in ((self.casl, L.F_CASL), (self.inferatt, L.F_INFERATT), (self.inferart, L.F_INFERART),
                        (self.incwm, L.F_INCWM), (self.inferwm, L.F_INFERWM), (self.infert1, L.F_INFERT1),
                        (self.artonly, L.F_ARTONLY)):
            if on:
                f |= bit
        return f

    def kernel_model(self, device_array=None):
        """-> (L.Model, keepalive list).  `device_array(np.ndarray) -> object with data_ptr()` uploads the
        per-voxel partial volumes when they are arrays."""
        m = L.Model()
        keep = []
        m.kind = self.KIND
        m.flags = self.kernel_flags()
        m.tau, m.t1b = self.tau, self.t1b
        m.t1, m.pc, m.fcalib, m.att = self.t1, self.pc, self.fcalib, self.att
        m.t1wm, m.pcwm, m.fcalibwm, m.attwm, m.fwm = self.t1wm, self.pcwm, self.fcalibwm, self.attwm, self.fwm
        m.artt, m.leadscale = self.artt, self.leadscale
        for name in ("pvgm", "pvwm"):
            value = getattr(self, name)
            if isinstance(value, np.ndarray):
                if device_array is None:
                    raise ValueError("per-voxel %s needs a device uploader" % name)
                buf = device_array(np.ascontiguousarray(value, dtype=np.float32))
                keep.append(buf)
                setattr(m, name, buf.data_ptr())
            else:
                setattr(m, name + "_s", float(value))
        return m, keep

    # ---- Model API ----
    def evaluate(self, params, tpts):
        """
        :param tpts: time values, shape [W, 1, N] or [1, 1, N] (also [n, N] with params [P, n, 1])
        :param params: sequence of P arrays [W, S, 1], or one [P, W, S, 1] array/tensor
        :return: [W, S, N] model output (a CUDA tensor)
        """
        from ..ops import evaluate_model
        n_params = len(params) if isinstance(params, (list, tuple)) else int(params.shape[0])
        if n_params != len(self.params):
            raise ValueError(f"Model set up to infer {len(self.params)} parameters; "
                             "this many parameter arrays must be supplied")
        return evaluate_model(self, params, tpts)

    def tpts(self):
        n_expected = len(self.tis) * self.repeats
        if self.data_model.n_tpts != n_expected:
            raise ValueError("ASL model configured with %i time points, but data has %i"
                             % (n_expected, self.data_model.n_tpts))
        base = np.repeat(np.asarray(self.tis, dtype=np.float64), self.repeats)       # grouped by TI/PLD
        z = self.data_model.voxel_coords()[:, 2].astype(np.float64)
        return (base[None, :] + (z * self.slicedt)[:, None]).astype(NP_DTYPE)

    def tpts_lowrank(self):
        """-> (ti [T] float32, zoff [W] float32 or None) with tpts()[w, r] == ti[r] + zoff[w] up to rounding."""
        ti = np.repeat(np.asarray(self.tis, dtype=np.float64), self.repeats).astype(NP_DTYPE)
        if not self.slicedt:
            return ti, None
        z = self.data_model.voxel_coords()[:, 2].astype(np.float64)
        return ti, (z * self.slicedt).astype(NP_DTYPE)

    def __str__(self):
        return "ASL resting state model: %s" % __version__

    # ---- posterior initialisers: f(param, t, data) -> (mean, var or None)  (aslrest.py:461-520) ----
    # `data` is [n, T] for the n voxels being initialised (the engine passes its own shard).  When the engine has
    # already reduced the data on the GPU (ops.InitData.device_stats, svbasl_init_stats) those statistics are used;
    # with a plain array the reductions run in numpy, as in the reference.
    @staticmethod
    def _stat(data, key, fallback):
        st = getattr(data, "device_stats", None)
        return np.asarray(st[key], dtype=NP_DTYPE) if st is not None else fallback(np.asarray(data)).astype(NP_DTYPE)

    def _shard_of(self, data, per_voxel):
        """Per-voxel option arrays cover all nodes; the engine initialises one shard at a time."""
        per_voxel = np.asarray(per_voxel, dtype=NP_DTYPE)
        sl = getattr(data, "voxel_slice", None)
        return per_voxel[sl] if (per_voxel.ndim and sl is not None) else per_voxel

    def _init_flow(self, _param, _t, data):
        f = np.maximum(self._stat(data, "mean_t", lambda d: d.mean(-1)), 0.1)
        if not self.pvcorr:
            return f, None
        # PVEc: assume GM:WM perfusion 3:1 (aslrest.py:470-483)
        fwm = f / (1 + 2 * self._shard_of(data, self.pvgm))
        return (fwm if _param.name == "fwm" else 3 * fwm), None

    def _init_fblood(self, _param, _t, data):
        return np.maximum(self._stat(data, "max_t", lambda d: d.max(axis=1)), 0.1), None

    def _init_delt(self, _param, t, data):
        n = len(data)
        if self.att_init == "max":
            def host_t_at_max(d):
                idx = np.argmax(d, axis=1)
                return np.take_along_axis(np.broadcast_to(np.asarray(t), d.shape), idx[:, None], axis=1)[:, 0]
            t_max = self._stat(data, "t_at_max", host_t_at_max)
            offset = 0.3 if _param.name == "fwm" else 0.0                # as written (never true for a delt)
            return (t_max + offset - self.tau).astype(NP_DTYPE), np.full(n, self.attsd, dtype=NP_DTYPE)
        # the reference returns attsd (a standard deviation) in the variance slot (aslrest.py:520)
        return np.full(n, self.att, dtype=NP_DTYPE), np.full(n, self.attsd, dtype=NP_DTYPE)
