"""bench.py's reference arm (the reference's own CPU path, timed on the host cores) runs without a GPU: check the
contract of its JSON line here; the GPU arm's line is checked on the GPU box."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

REQUIRED = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e")


def _run(extra_env=None, *args):
    env = dict(os.environ, **(extra_env or {}))
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", *args],
                          capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_prints_one_json_line(oracle_mod):
    r = _run()
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.split("\n") if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in REQUIRED:
        assert k in d, k
    assert d["impl"] == "reference" and d["vs_baseline"] is None and d["higher_is_better"] is True
    assert d["unit"] == "correlations/s" and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2")
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.gpu
def test_gpu_arm_line_has_the_contract_keys():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "3", "--warmup", "3", "--no-grid", "--no-cpu-baseline"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.split("\n") if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in REQUIRED:
        if k != "cpu_baseline":
            assert k in d, k
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert d["gpu_launches"] == 3 * rf["launches_per_step"] * d["steps"] and d["e2e"]["h2d_bytes_per_step"] > 0 and d["clocks"]["sm_mhz"] > 0
    assert 0 < rf["fp32_frac"] < 1 and rf["fp32_peak_tflops"] == 74.4 and rf["traffic"] > 0
    assert d["ms_per_step"] * d["steps"] >= 150            # 3 steps of ~60 ms: the default 40 steps run for seconds
