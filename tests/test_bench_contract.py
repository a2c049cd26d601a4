"""The bench line contract (task statement section 4): the committed lines under profiles/ -- written by bench.py on the
GPU box -- carry every required key with sane values; `bench.py --impl reference` is exercised for real (CPU only)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "e2e", "gpu_launches")


def _load(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        lines = [ln for ln in f.read().splitlines() if ln.strip()]
    assert len(lines) == 1, "exactly one JSON line"
    return json.loads(lines[0])


def test_own_arm_lines_carry_the_contract():
    for name, n in (("r01_bench_line.json", 1), ("r01_bench_line_2gpu.json", 2), ("r01_bench_line_8gpu.json", 8)):
        d = _load(name)
        for k in BASE + ("roofline", "clocks"):
            assert k in d, (name, k)
        assert d["n_gpus"] == n and d["unit"] == "nnz/s" and d["dtype"] == "f64" and d["scaling"] == "weak" and d["higher_is_better"] is True
        assert d["vs_baseline"] is None and "workload" in d["config"] and "model" not in d["config"]
        assert abs(d["value"] - n * 89_999_985 / (d["ms_per_step"] * 1e-3)) <= 0.02 * d["value"]          # whole-job aggregate
        r = d["roofline"]
        assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.5 < r["frac"] < 1.0
        assert r["traffic"] is None or 0.8 < r["traffic"] / r["algorithmic_bytes_per_launch"] < 1.2
        e = d["e2e"]
        assert e["unit"] == "nnz/s" and 0 < e["value"] < d["value"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] >= 8 * 89_999_985 * n
        assert d["gpu_launches"] == d["steps"]                                                           # one generated kernel per step
        assert d["clocks"]["sm_max_mhz"] > 0 and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    one = _load("r01_bench_line.json")
    c = one["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["unit"] == "nnz/s" and c["value"] > 0 and "sample" in c


def test_reference_arm_runs_here_and_prints_one_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in BASE + ("cpu_baseline", "impl"):
        assert k in d, k
    assert d["impl"] == "reference" and d["gpu_launches"] == 0 and d["cpu_baseline"]["kind"] in ("port", "port-compiled")
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 1e6
