"""
CPU tier: host-side logic of the C ABI that runs before any CUDA call - kernel-table coverage of every plugin
option combination, descriptor validation and its error messages (error behaviour mirrors the reference's
ValueError texts where it has them: aslrest.py:97-120, 248-262).  No compute is launched.
"""
import ctypes as C
import itertools

import numpy as np
import pytest

from svb import DataModel
from svb_models_asl_b200 import _lib as L
from svb_models_asl_b200.plugin import get_model_class

PLDS = [0.25, 0.5, 0.75, 1.0, 1.25, 1.5]


@pytest.fixture(scope="module")
def lib():
    return L.load()


def _dm(n=4, t=6):
    return DataModel(np.zeros((n, t), dtype=np.float32))


def _all_aslrest_options():
    for casl, att, art, t1 in itertools.product((False, True), repeat=4):
        for wm in ({}, {"incwm": True, "pvgm": 0.6, "pvwm": 0.4}, {"incwm": True, "inferwm": True, "pvgm": 0.6, "pvwm": 0.4},
                   {"inferwm": True}):                 # WM parameters without a WM signal (aslrest.py:197-211 vs :327)
            yield dict(casl=casl, inferatt=att, inferart=art, infert1=t1, **wm)
    for casl, att in itertools.product((False, True), repeat=2):
        yield dict(casl=casl, inferatt=att, inferart=True, artonly=True)


def test_every_aslrest_option_combination_has_a_kernel(lib):
    """len(model.params) (aslrest.py:183-246) == P of the kernel the C ABI dispatches to, for all 68 layouts."""
    cls = get_model_class("aslrest")
    seen = set()
    for opts in _all_aslrest_options():
        model = cls(_dm(), tau=1.8, plds=PLDS, **opts)
        m, _keep = model.kernel_model()
        p = lib.svbasl_model_n_params(C.byref(m))
        assert p == len(model.params), (opts, p, [q.name for q in model.params])
        seen.add(m.flags)
    assert len(seen) == 68


def test_every_disp_option_combination_has_a_kernel(lib):
    cls = get_model_class("aslrest_disp")
    n = 0
    for casl, att, art, infer in itertools.product((False, True), repeat=4):
        for artonly in ((False, True) if art else (False,)):
            model = cls(_dm(), tau=1.8, plds=PLDS, casl=casl, inferatt=att, inferart=art, artonly=artonly,
                        disptype="gamma", inferdisp=infer)
            m, _keep = model.kernel_model()
            assert lib.svbasl_model_n_params(C.byref(m)) == len(model.params)
            n += 1
    assert n == 24


def test_unknown_layout_is_reported_not_guessed(lib):
    m = L.Model()
    m.kind, m.flags = 7, 0             # not a model family
    assert lib.svbasl_model_n_params(C.byref(m)) < 0
    assert b"no kernel compiled for model kind=7" in lib.svbasl_last_error()
    assert lib.svbasl_model_n_params(None) < 0


def _engine(n_par=3, n_vox=8, **kw):
    e = L.Engine()
    e.n_vox, e.w_begin, e.ld = n_vox, 0, n_vox
    e.n_vox_global = n_vox
    e.n_par, e.n_samples, e.n_batch, e.t_full = n_par, 5, 6, 6
    e.t_row_stride = 1
    e.latent = L.LATENT_NUMERIC
    for i in range(n_par):
        e.prior_type[i] = L.PRIOR_CODES["N"]
        e.prior_var[i] = 1.0
    e.latent_weight, e.grad_scale = 1.0, 1.0
    # never dereferenced: validation fails first in every case below
    e.state = e.data = e.tpts = 0x1000
    for k, v in kw.items():
        setattr(e, k, v)
    return e


def _model(lib):
    model = get_model_class("aslrest")(_dm(), tau=1.8, casl=True, plds=PLDS)
    return model.kernel_model()[0]


@pytest.mark.parametrize("patch,needle", [
    (dict(n_vox=-1), b"bad extents"),
    (dict(ld=4), b"bad extents"),
    (dict(n_samples=0), b"bad sizes"),
    (dict(t_row_stride=0), b"bad sizes"),
    (dict(state=None), b"state, data and tpts|ti are required"),
    (dict(tpts=None, ti=None), b"state, data and tpts|ti are required"),
    (dict(n_par=4), b"engine n_par=4 but the model has 2 parameters"),
])
def test_step_rejects_bad_descriptors_before_touching_the_gpu(lib, patch, needle):
    m = _model(lib)
    e = _engine(**patch)
    if "n_par" in patch:
        for i in range(patch["n_par"]):
            e.prior_type[i] = L.PRIOR_CODES["N"]
    cost = C.c_void_p(0x1000)
    rc = lib.svbasl_elbo_grad(C.byref(m), C.byref(e), 0, cost, None, None, None)
    assert rc < 0
    assert needle in lib.svbasl_last_error(), lib.svbasl_last_error()


def test_spatial_prior_requirements_are_checked(lib):
    m = _model(lib)
    e = _engine()
    e.prior_type[0] = L.PRIOR_CODES["M"]
    assert lib.svbasl_elbo_grad(C.byref(m), C.byref(e), 0, None, None, None, None) < 0
    assert b"spatial prior without neighbours" in lib.svbasl_last_error()
    e.latent = L.LATENT_ANALYTIC
    assert lib.svbasl_elbo_grad(C.byref(m), C.byref(e), 0, None, None, None, None) < 0
    assert b"sample-based latent loss" in lib.svbasl_last_error()
    # an update with a spatial prior is one iteration per launch
    e.latent = L.LATENT_NUMERIC
    e.neighbours = e.log_ak = e.spatial_samples = 0x1000
    ad = L.Adam()
    ad.m = ad.v = ad.lr_t = 0x1000
    ad.n_iters, ad.n_batches = 2, 1
    assert lib.svbasl_step(C.byref(m), C.byref(e), C.byref(ad), None, None, None) < 0
    assert b"n_iters must be 1" in lib.svbasl_last_error()
    # next-iteration samples: a second buffer and the in-kernel draws
    ad.n_iters = 1
    e.spatial_samples_out = e.spatial_samples
    assert lib.svbasl_step(C.byref(m), C.byref(e), C.byref(ad), None, None, None) < 0
    assert b"spatial_samples_out needs" in lib.svbasl_last_error()
    e.spatial_samples_out = 0x2000
    e.eps = 0x3000
    assert lib.svbasl_step(C.byref(m), C.byref(e), C.byref(ad), None, None, None) < 0
    assert b"spatial_samples_out needs" in lib.svbasl_last_error()
    e.eps = None
    # peer pointers must name the first / last owned voxels of the launch
    e.peer_lo, e.peer_lo_first, e.peer_lo_count = 0x4000, e.w_begin + 1, 2
    assert lib.svbasl_step(C.byref(m), C.byref(e), C.byref(ad), None, None, None) < 0
    assert b"first / last owned voxels" in lib.svbasl_last_error()
    e.peer_lo = None
    # fused hyper tail: needs the device counter shared with the engine
    hy = L.Hyper()
    hy.log_ak, hy.m, hy.v, hy.lr_t, hy.done_ctas, hy.step_dev = e.log_ak, 0x1000, 0x1000, 0x1000, 0x1000, 0x5000
    hy.n_spatial, hy.world = 1, 1
    e.ak_grad = 0x1000
    assert lib.svbasl_step_spatial(C.byref(m), C.byref(e), C.byref(ad), C.byref(hy), None, None, None) < 0
    assert b"bad svbasl_hyper descriptor" in lib.svbasl_last_error()
    e.step_dev = 0x5000
    hy.world, hy.status = 2, 0x1000
    hy.mailboxes[0] = 0x6000
    assert lib.svbasl_step_spatial(C.byref(m), C.byref(e), C.byref(ad), C.byref(hy), None, None, None) < 0
    assert b"mailbox of rank 1 is NULL" in lib.svbasl_last_error()


def test_null_and_range_checks_of_the_small_entry_points(lib):
    assert lib.svbasl_step(None, None, None, None, None, None) < 0
    assert lib.svbasl_hyper_step(None, None, None, None, 1, 1.0, 0.1, 0.9, 0.999, 1e-8, None) < 0
    assert lib.svbasl_hyper_step_dev(None, None, None, None, 1, 1.0, None, None, 0.9, 0.999, 1e-8, None) < 0
    boxes = (C.c_void_p * 2)(0x1000, None)
    one = C.c_void_p(0x1000)
    rc = lib.svbasl_hyper_step_peers(one, one, one, one, 1, 1.0, one, one, 0.9, 0.999, 1e-8, 0, 2, boxes, one, None)
    assert rc < 0 and b"mailbox of rank 1 is NULL" in lib.svbasl_last_error()
    rc = lib.svbasl_hyper_step_peers(one, one, one, one, 1, 1.0, one, one, 0.9, 0.999, 1e-8, 0, 99, boxes, one, None)
    assert rc < 0 and b"bad hyper_step_peers arguments" in lib.svbasl_last_error()
    assert lib.svbasl_mailbox_bytes(8) == 2 * 8 * 40 and lib.svbasl_mailbox_bytes(0) == 0
    assert lib.svbasl_shared_alloc(0, None, None) < 0
    assert lib.svbasl_sample_spatial(None, 0, 0, None, None) < 0
    assert lib.svbasl_sample_spatial_next(None, 0, None, None) < 0
    assert lib.svbasl_abi_version() == 2


def test_plugin_option_errors_match_the_reference_texts():
    """aslrest.py:97-120: the option checks a user of the reference relies on."""
    cls = get_model_class("aslrest")
    with pytest.raises(ValueError, match="Either TIs or PLDs"):
        cls(_dm(), tau=1.8)
    with pytest.raises(NotImplementedError, match="Variable repeats"):
        cls(_dm(), tau=1.8, plds=PLDS, repeats=[1, 2])
    model = cls(_dm(n=4, t=12), tau=1.8, plds=PLDS, repeats=[2])
    assert model.tpts().shape[-1] == 12
    with pytest.raises(ValueError, match="time points"):
        cls(_dm(n=4, t=7), tau=1.8, plds=PLDS).tpts()


def test_empty_shard_is_a_no_op(lib):
    """n_vox == 0 (a rank that owns nothing): every entry point returns 0 without launching anything."""
    m = _model(lib)
    e = _engine(n_vox=0, ld=8)
    assert lib.svbasl_elbo_grad(C.byref(m), C.byref(e), 0, None, None, None, None) == 0
    ad = L.Adam()
    ad.m = ad.v = ad.lr_t = 0x1000
    ad.n_iters, ad.n_batches = 1, 1
    ad.beta1, ad.beta2, ad.epsilon = 0.9, 0.999, 1e-8
    assert lib.svbasl_step(C.byref(m), C.byref(e), C.byref(ad), None, None, None) == 0
    assert lib.svbasl_sample_spatial(C.byref(e), 0, 0, C.c_void_p(0x1000), None) == 0
