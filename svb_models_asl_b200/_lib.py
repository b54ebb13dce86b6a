"""
ctypes binding of libsvbasl.so (include/svbasl.h).  Thin on purpose: structures mirror the C ABI field for
field, every call checks the return code and raises with svbasl_last_error().

There is NO fallback: if the CUDA library is missing or cannot be loaded, importing the ops raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# SVBASL_LIB: load another build of the same library (e.g. a tuning variant); default = the in-tree build
LIB_PATH = os.environ.get("SVBASL_LIB") or os.path.join(HERE, "csrc", "libsvbasl.so")

MAX_PAR = 10
MAX_SPATIAL = 4
MAX_PEERS = 16

MODEL_ASLREST, MODEL_ASLREST_DISP, MODEL_ASLNN = 0, 1, 2
F_CASL, F_INFERATT, F_INFERART, F_INCWM, F_INFERWM, F_INFERT1, F_ARTONLY, F_DISP_INFER, F_DISP_ASWRITTEN, F_NN_TC = (
    0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40, 0x80, 0x100, 0x200)
XF_IDENTITY, XF_EXP, XF_ABS = 0, 1, 2
PRIOR_N, PRIOR_ARD, PRIOR_MRF = 0, 1, 2
LATENT_NUMERIC, LATENT_ANALYTIC = 0, 1
PRIOR_CODES = {"N": PRIOR_N, "A": PRIOR_ARD, "M": PRIOR_MRF}

c_float_p = C.POINTER(C.c_float)


class Model(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("flags", C.c_uint32),
        ("tau", C.c_float), ("t1b", C.c_float),
        ("t1", C.c_float), ("pc", C.c_float), ("fcalib", C.c_float), ("att", C.c_float),
        ("t1wm", C.c_float), ("pcwm", C.c_float), ("fcalibwm", C.c_float), ("attwm", C.c_float), ("fwm", C.c_float),
        ("artt", C.c_float), ("leadscale", C.c_float),
        ("pvgm_s", C.c_float), ("pvwm_s", C.c_float),
        ("pvgm", C.c_void_p), ("pvwm", C.c_void_p),
        ("conv_dt", C.c_float), ("conv_tmax", C.c_float), ("conv_nt", C.c_int32),
        ("s_fixed", C.c_float), ("sp_fixed", C.c_float),
        ("nn_weights", C.c_void_p),
    ]


class Engine(C.Structure):
    _fields_ = [
        ("n_vox", C.c_int64), ("w_begin", C.c_int64), ("ld", C.c_int64), ("vox_offset", C.c_int64),
        ("n_vox_global", C.c_int64),
        ("n_par", C.c_int32), ("n_samples", C.c_int32), ("n_batch", C.c_int32), ("t_full", C.c_int32),
        ("latent", C.c_int32), ("cov_llt", C.c_int32),
        ("prior_type", C.c_int32 * MAX_PAR), ("prior_mean", C.c_float * MAX_PAR), ("prior_var", C.c_float * MAX_PAR),
        ("ard_phi_max", C.c_float), ("latent_weight", C.c_float), ("grad_scale", C.c_float),
        ("state", C.c_void_p), ("state_out", C.c_void_p),
        ("data", C.c_void_p), ("tpts", C.c_void_p), ("ti", C.c_void_p), ("zoff", C.c_void_p),
        ("t_row0", C.c_int32), ("t_row_stride", C.c_int32),
        ("eps", C.c_void_p), ("seed", C.c_uint64),
        ("neighbours", C.c_void_p), ("spatial_samples", C.c_void_p), ("spatial_samples_out", C.c_void_p),
        ("log_ak", C.c_void_p), ("ak_grad", C.c_void_p),
        ("step_dev", C.c_void_p), ("cost_sum_scalar", C.c_int32),
        ("peer_lo", C.c_void_p), ("peer_hi", C.c_void_p),
        ("peer_lo_ld", C.c_int64), ("peer_hi_ld", C.c_int64), ("peer_lo_shift", C.c_int64), ("peer_hi_shift", C.c_int64),
        ("peer_lo_first", C.c_int64), ("peer_lo_count", C.c_int64), ("peer_hi_first", C.c_int64), ("peer_hi_count", C.c_int64),
    ]


class Adam(C.Structure):
    _fields_ = [
        ("m", C.c_void_p), ("v", C.c_void_p), ("lr_t", C.c_void_p),
        ("beta1", C.c_float), ("beta2", C.c_float), ("epsilon", C.c_float),
        ("step0", C.c_int64), ("n_iters", C.c_int32), ("n_batches", C.c_int32),
    ]


class Hyper(C.Structure):
    _fields_ = [
        ("log_ak", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("lr_t", C.c_void_p), ("step_dev", C.c_void_p),
        ("done_ctas", C.c_void_p),
        ("beta1", C.c_float), ("beta2", C.c_float), ("epsilon", C.c_float),
        ("n_spatial", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32),
        ("mailboxes", C.c_void_p * MAX_PEERS), ("status", C.c_void_p),
    ]


class SvbAslError(RuntimeError):
    pass


_EXPORTS = {
    "svbasl_last_error": (C.c_char_p, []),
    "svbasl_abi_version": (C.c_int, []),
    "svbasl_model_n_params": (C.c_int, [C.POINTER(Model)]),
    "svbasl_n_state": (C.c_int, [C.POINTER(Model), C.POINTER(Engine)]),
    "svbasl_evaluate": (C.c_int, [C.POINTER(Model), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                  C.c_int32, C.c_int64, C.c_void_p]),
    "svbasl_nn_pack_weights": (C.c_int, [C.POINTER(Model), C.c_void_p]),
    "svbasl_nn_evaluate_tc": (C.c_int, [C.POINTER(Model), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
    "svbasl_elbo_grad": (C.c_int, [C.POINTER(Model), C.POINTER(Engine), C.c_int64, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p]),
    "svbasl_step": (C.c_int, [C.POINTER(Model), C.POINTER(Engine), C.POINTER(Adam), C.c_void_p, C.c_void_p,
                              C.c_void_p]),
    "svbasl_step_spatial": (C.c_int, [C.POINTER(Model), C.POINTER(Engine), C.POINTER(Adam), C.POINTER(Hyper), C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "svbasl_sample_spatial": (C.c_int, [C.POINTER(Engine), C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "svbasl_sample_spatial_next": (C.c_int, [C.POINTER(Engine), C.c_int64, C.c_void_p, C.c_void_p]),
    "svbasl_hyper_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_float,
                                    C.c_float, C.c_float, C.c_float, C.c_void_p]),
    "svbasl_hyper_step_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_void_p,
                                        C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p]),
    "svbasl_mailbox_bytes": (C.c_int64, [C.c_int32]),
    "svbasl_hyper_step_peers": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_void_p,
                                          C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int32, C.c_int32,
                                          C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p]),
    "svbasl_shared_alloc": (C.c_int, [C.c_int64, C.POINTER(C.c_void_p), C.c_void_p]),
    "svbasl_shared_free": (C.c_int, [C.c_void_p]),
    "svbasl_shared_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "svbasl_shared_close": (C.c_int, [C.c_void_p]),
    "svbasl_advance_step": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p]),
    "svbasl_fill_eps": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_uint64,
                                  C.c_int64, C.c_void_p]),
    "svbasl_init_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "svbasl_model_fit": (C.c_int, [C.POINTER(Model), C.POINTER(Engine), C.c_void_p, C.c_void_p]),
    "svbasl_host_ctx_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64, C.c_int32]),
    "svbasl_host_ctx_destroy": (C.c_int, [C.c_void_p]),
    "svbasl_step_host": (C.c_int, [C.c_void_p, C.POINTER(Model), C.POINTER(Engine), C.POINTER(Adam), C.POINTER(Hyper),
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "svbasl_host_sync": (C.c_int, [C.c_void_p]),
}

_lib = None


def exported_symbols():
    """Names include/svbasl.h declares (the not-gpu test checks the library exports each one)."""
    return sorted(_EXPORTS)


def load():
    """Load libsvbasl.so (once).  Raises SvbAslError if it has not been built - there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SvbAslError("libsvbasl.so not found at %s: build it with `python -m svb_models_asl_b200.build` "
                          "(CUDA 12.9 nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as exc:
        raise SvbAslError("cannot load %s: %s" % (LIB_PATH, exc)) from exc
    for name, (res, args) in _EXPORTS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.svbasl_abi_version() != 2:
        raise SvbAslError("libsvbasl.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc):
    if rc < 0:
        raise SvbAslError("svbasl error %d: %s" % (rc, load().svbasl_last_error().decode()))
    return rc
