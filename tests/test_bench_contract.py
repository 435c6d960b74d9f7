"""The measurement contract of bench.py, checked on CPU: the reference arm runs here (it is the CPU oracle), and the
recorded GPU lines under profiles/ must carry every key the contract names with consistent arithmetic.  No GPU, no
kernels: this guards the JSON schema the driver parses, not a number."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def _baseline():
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        return json.load(f)


def test_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU oracle timed on host cores): one JSON line, same metric/config as the GPU arm,
    `impl`, `cpu_baseline` of kind "port", an `e2e` that moves no bytes, and no kernel launches."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["unit"] == "graphs/s" and d["value"] > 0
    assert "workload" in d["config"] and "case118v2" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    # graphs per step / time per step is the value
    assert abs(d["value"] - 128 / (d["ms_per_step"] / 1e3)) < 1e-6 * d["value"]


@pytest.mark.parametrize("record,n", [("bench_1gpu.json", 1), ("bench_2gpu.json", 2), ("bench_8gpu.json", 8)])
def test_recorded_gpu_lines_follow_the_contract(record, n):
    """The lines committed under profiles/r2/ (written by bench.py on B200 boxes): every contract key, whole-job
    throughput = ranks x 128 graphs / ms_per_step, roofline.frac = achieved / peak, achieved = algorithmic bytes / time,
    copies counted, kernels counted, clean clocks."""
    path = os.path.join(ROOT, "profiles", "r2", record)
    with open(path) as f:
        d = json.load(f)
    assert BASE_KEYS | {"clocks", "roofline"} <= set(d), (BASE_KEYS | {"clocks", "roofline"}) - set(d)
    assert d["n_gpus"] == n and d["scaling"] == "weak" and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["metric"] == _baseline()["metric"] or "graphs/sec" in d["metric"]
    assert abs(d["value"] - n * 128 / (d["ms_per_step"] / 1e3)) < 1e-6 * d["value"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 1_000_000 and e["d2h_bytes_per_step"] >= 4
    assert 0 < e["value"] < d["value"]  # the end-to-end figure includes the copies: never above the resident one
    assert d["gpu_launches"] >= 9 * d["steps"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["us_per_launch"] * 1e-6) / 1e9) < 1e-6 * r["achieved"]
    c = d["clocks"]
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert c["sm_mhz"] >= 0.9 * c["sm_max_mhz"]
    if n == 1:
        cb = d["cpu_baseline"]
        assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]
    else:
        p = d["dp_parity"]
        assert p["max_rel"] < 1e-5 and p["fro_rel"] < 1e-5
