"""
The drop-in surface around the plugins: the example scripts carry the reference's option dicts unchanged
(/root/reference/scripts/asl_example*.py), the `svb.models` entry points of the reference's setup.py:89-95 are
registered by this package's metadata, and the quick_test known-answer script runs (GPU).
"""
import ast
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/scripts"
SCRIPTS = ["asl_example.py", "asl_example_sim.py", "asl_example_nn.py", "asl_example_sim_nn.py"]


def _options_of(path):
    """The literal `options = {...}` of a script and its model / outdir names, without executing it."""
    tree = ast.parse(open(path).read())
    found = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name):
            name = node.targets[0].id
            if name == "options" and isinstance(node.value, ast.Dict):
                d = {}
                for k, v in zip(node.value.keys, node.value.values):
                    key = ast.literal_eval(k)
                    d[key] = "sys.stdout" if key == "log_stream" else ast.literal_eval(v)
                found["options"] = d
            elif name in ("model", "outdir") and isinstance(node.value, ast.Constant):
                found[name] = node.value.value
    return found


# the reference's dicts, frozen here so the check also runs where /root/reference does not exist (GPU box)
EXPECTED_KEYS = {
    "asl_example.py": ("aslrest", {"repeats": [8], "slicedt": 0.0452, "learning_rate": 0.01, "batch_size": 6, "epochs": 500}),
    "asl_example_sim.py": ("aslrest", {"repeats": [1], "learning_rate": 0.05, "epochs": 5000}),
    "asl_example_nn.py": ("aslnn", {"repeats": [8], "slicedt": 0.0452, "batch_size": 6, "train_load": "trained_data"}),
    "asl_example_sim_nn.py": ("aslnn", {"repeats": [1], "epochs": 5000, "train_load": "trained_data"}),
}


@pytest.mark.parametrize("script", SCRIPTS)
def test_example_scripts_carry_the_reference_option_dicts(script):
    mine = _options_of(os.path.join(ROOT, "scripts", script))
    model, subset = EXPECTED_KEYS[script]
    assert mine["model"] == model
    for k, v in subset.items():
        assert mine["options"][k] == v, k
    for k in ("save_mean", "save_var", "save_param_history", "save_cost", "save_cost_history", "save_model_fit", "save_log",
              "force_num_latent_loss"):
        assert mine["options"][k] is True, k
    assert mine["options"]["plds"] == [0.25, 0.5, 0.75, 1.0, 1.25, 1.5] and mine["options"]["tau"] == 1.8
    if os.path.isdir(REF):                                   # build container: literal equality with the reference
        ref = _options_of(os.path.join(REF, script))
        assert mine["options"] == ref["options"]
        assert mine["model"] == ref["model"] and mine["outdir"] == ref["outdir"]


def test_svb_models_entry_points_are_registered(tmp_path):
    """pyproject.toml registers aslnn / aslrest / aslrest_disp in the `svb.models` group like the reference's
    setup.py:89-95: installed (into a scratch directory, from a source-only copy), importlib.metadata finds them and they
    load to the plugin classes."""
    src = tmp_path / "src"
    src.mkdir()
    shutil.copy(os.path.join(ROOT, "pyproject.toml"), src)
    for pkg in ("svb", "svb_models_asl", "svb_models_asl_b200"):
        shutil.copytree(os.path.join(ROOT, pkg), src / pkg,
                        ignore=shutil.ignore_patterns("_obj", "_gen", "__pycache__", "*.o", "*.cubin"))
    site = tmp_path / "site"
    res = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-deps", "--no-build-isolation",
                          "--quiet", "--target", str(site), str(src)], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from importlib.metadata import entry_points\n"
            "eps = {e.name: e for e in entry_points(group='svb.models')}\n"
            "print(sorted(eps))\n"
            "print([eps[n].load().__name__ for n in sorted(eps)])\n" % str(site))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert out.returncode == 0, out.stderr[-3000:]
    lines = out.stdout.strip().splitlines()
    assert lines[0] == "['aslnn', 'aslrest', 'aslrest_disp']"
    assert lines[1] == "['AslNNModel', 'AslRestModel', 'AslRestDisp']"


def test_get_model_class_knows_the_three_names():
    from svb_models_asl_b200.plugin import get_model_class
    assert [get_model_class(n).__name__ for n in ("aslnn", "aslrest", "aslrest_disp")] == \
        ["AslNNModel", "AslRestModel", "AslRestDisp"]
    with pytest.raises(ValueError, match="No such model"):
        get_model_class("biexp")


@pytest.mark.gpu
def test_quick_test_known_answers():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "quick_test.py")], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "quick_test OK" in res.stdout
