"""
CPU tier: the parts of the bench.py contract that do not need a GPU - the reference arm (`--impl reference`: the
CPU port of the reference graph, one JSON line with the agreed keys) and the committed roofline inputs.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--voxels", "500"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["unit"] == "voxel-iters/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    # the reference arm runs the configuration it is asked for (the driver asks for none: the GPU arm's 1M voxels)
    assert d["config"]["voxels_per_gpu"] == 500 and d["config"]["name"] == "sim_art"


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "1", "--voxels", "500"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_committed_roofline_inputs_are_consistent():
    traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    # measured DRAM bytes per voxel of the headline kernel stay below the algorithmic 8B + 24 n_state = 552
    assert 300 < traffic["sim_art"]["dram_bytes_per_voxel"] <= 552
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert "north_star" in base and len(base["configs"]) >= 2
