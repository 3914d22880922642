"""The bench line contract (driver-facing): every committed bench line under profiles/ of the current round carries the
keys the contract names, with consistent values.  CPU only: checks the stored lines, not a run."""
import glob
import json
import os

import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"]


def _lines():
    return sorted(p for p in glob.glob(os.path.join(ROOT, "profiles", "r01d_bench_*.json")) if "reference" not in p)


@pytest.mark.parametrize("path", _lines(), ids=os.path.basename)
def test_bench_line_has_contract_keys(path):
    d = json.loads(open(path).read().strip().splitlines()[-1])
    for k in REQUIRED:
        assert k in d, k
    assert d["metric"] == "images/sec" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["data"] == "synthetic" and d["dtype"] == "bf16"
    assert "workload" in d["config"] and "model" not in d["config"]
    images = d["config"]["global_batch"] * d["steps"]
    assert abs(d["value"] - images / (d["ms_per_step"] * d["steps"] * 1e-3)) < 1e-6 * d["value"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < 1.15 * d["value"]
    r = d["roofline"]
    assert r["bound"] == "tensor" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
    assert d["gpu_launches"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if d["n_gpus"] == 1 and "cpu_baseline" in d:
        assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1


def test_reference_arm_line():
    path = os.path.join(ROOT, "profiles", "r01d_bench_reference_arm.json")
    d = json.loads(open(path).read().strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "images/sec" and d["unit"] == "images/s"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] == d["cpu_baseline"]["value"]
