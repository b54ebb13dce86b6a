"""
CPU tier: the plugin classes a user of the reference imports (svb_models_asl.AslRestModel / AslRestDisp / AslNNModel,
here aliases of svb_models_asl_b200.plugin) expose the same parameter lists, prior / posterior distributions and
derived constants as the reference's own classes, for the option sets of the golden fixtures
(tests/golden/make_golden.py ran the reference source: aslrest.py:69-246, aslrest_disp.py:30-38, aslnn.py:60-120).
"""
import numpy as np
import pytest

from svb import DataModel
from svb_models_asl import AslRestModel
from svb_models_asl_b200.plugin import get_model_class

TIS = [2.05, 2.3, 2.55, 2.8, 3.05, 3.3]

CASES = {
    "casl_tiss": dict(casl=True),
    "pasl_tiss": dict(casl=False),
    "casl_tiss_art": dict(casl=True, inferart=True),
    "pasl_tiss_art": dict(casl=False, inferart=True),
    "casl_noatt": dict(casl=True, inferatt=False),
    "casl_artonly": dict(casl=True, artonly=True),
    "casl_t1": dict(casl=True, infert1=True),
    "pasl_t1_art": dict(casl=False, infert1=True, inferart=True),
    "casl_pvc": dict(casl=True, pvcorr=True, inferart=True),
    "casl_pvc_t1": dict(casl=True, pvcorr=True, infert1=True),
    "casl_inferwm_noinc": dict(casl=True, inferwm=True, inferart=True),
}


def _meta(model):
    ps = model.params
    return {"names": [p.name for p in ps],
            "prior_types": [p.prior_type for p in ps],
            "prior_mean": np.asarray([np.mean(p.prior_dist.mean) for p in ps], dtype=np.float32),
            "prior_var": np.asarray([np.mean(p.prior_dist.var) for p in ps], dtype=np.float32),
            "post_mean": np.asarray([np.mean(p.post_dist.mean) for p in ps], dtype=np.float32),
            "post_var": np.asarray([np.mean(p.post_dist.var) for p in ps], dtype=np.float32)}


@pytest.mark.parametrize("case", sorted(CASES))
def test_aslrest_parameters_and_distributions_equal_the_reference(golden, case):
    g = golden("aslrest_eval")[case]
    W = g["params"].shape[1]
    opts = dict(CASES[case])
    for k in ("pvgm", "pvwm"):
        if k in g:
            opts[k] = g[k]
    model = AslRestModel(DataModel(np.zeros((W, len(TIS)), dtype=np.float32)), tis=TIS, tau=1.8, t1b=1.65, repeats=1,
                         **opts)
    m = _meta(model)
    assert m["names"] == list(g["names"])
    assert m["prior_types"] == list(g["prior_types"])
    for k in ("prior_mean", "prior_var", "post_mean", "post_var"):
        np.testing.assert_allclose(m[k], g[k], rtol=1e-6, err_msg=k)
    assert float(np.mean(model.pc)) == pytest.approx(float(g["pc"]), rel=1e-6)
    assert float(model.attsd) == pytest.approx(float(g["attsd"]), rel=1e-6)
    assert float(model.artt) == pytest.approx(float(g["artt"]), rel=1e-6)


def test_disp_parameter_list_and_grid_equal_the_reference(golden):
    d = golden("disp_pieces")["disp"]
    cls = get_model_class("aslrest_disp")
    model = cls(DataModel(np.zeros((4, len(TIS)), dtype=np.float32)), tis=TIS, tau=1.8, t1b=1.65, repeats=1, casl=True,
                inferart=True, disptype="gamma", inferdisp=True)
    m = _meta(model)
    assert m["names"] == list(d["names"])
    np.testing.assert_allclose(m["prior_mean"], d["prior_mean"], rtol=1e-6)
    np.testing.assert_allclose(m["prior_var"], d["prior_var"], rtol=1e-6)
    assert model.conv_nt == int(d["conv_nt"]) and float(model.conv_tmax) == pytest.approx(float(d["conv_tmax"]))


def test_aslnn_parameter_list_and_default_time_points_equal_the_reference(golden, tmp_path):
    g = golden("aslnn_eval")["nn"]
    for i in range(3):
        np.save(tmp_path / ("weights%i.npy" % i), g["w%i" % i])
        np.save(tmp_path / ("biases%i.npy" % i), g["b%i" % i])
    cls = get_model_class("aslnn")
    model = cls(DataModel(np.zeros((1, 6), dtype=np.float32)), tis=TIS, tau=1.8, casl=True, train_load=str(tmp_path))
    assert [p.name for p in model.params] == list(g["names"])
    assert [type(p.post_dist).__name__ for p in model.params] == list(g["dists"])
    np.testing.assert_allclose(np.asarray(model.tpts()), g["tpts_default"], rtol=1e-6)
