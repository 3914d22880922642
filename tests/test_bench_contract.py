"""bench.py contract, exercised by RUNNING its CPU-side code (no GPU): the `--impl reference` arm end to end on the small
DLA-34 workload (one JSON line with the keys and invariants the driver reads), the non-zero-rank no-op, and the host helpers
the GPU arm relies on (usable-core detection, clock-sample parsing, the BASELINE.md-3 batch bounding of the CPU arm)."""
import json
import os
import subprocess
import sys

import bench

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def _run(extra, env=None):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "dla34",
                          "--steps", "1", "--warmup", "1"] + extra, capture_output=True, text=True, timeout=900,
                         env=dict(os.environ, **(env or {})))
    assert res.returncode == 0, res.stderr[-3000:]
    return [l for l in res.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_prints_one_contract_line():
    lines = _run([])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "images/sec" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["data"] == "synthetic" and d["dtype"] == "f32"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "min(B, 8)" in cb["sample"]
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] == 1 and d["warmup"] == 1


def test_reference_arm_other_ranks_exit_silently():
    assert _run(["--gpus", "2"], env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}) == []


def test_host_helpers():
    assert 1 <= bench.usable_cpus() <= (os.cpu_count() or 1)
    s = bench.ClockSampler(0)
    s.proc = object.__new__(subprocess.Popen)  # parsing only: no process behind it
    s.proc.terminate = lambda: None
    s.rows = [["1590", "1965", "Not Active", "Not Active", "Not Active", "Active"],
              ["1605", "1965", "Not Active", "Not Active", "Not Active", "Not Active"], ["garbage"]]
    out = s.stop()
    assert out["sm_mhz"] == 1597.5 and out["sm_max_mhz"] == 1965.0 and out["reasons"] == ["sw_power_cap"]
    assert set(bench.WORKLOADS) >= {"v2_99", "dla34"} and bench.WORKLOADS["v2_99"][2:5] == (32, 900, 1600)
    assert bench.load_peaks()["tflops"] > 100 and bench.load_peaks()["gbs"] > 1000
