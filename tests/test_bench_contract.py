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


def test_round_2_lines_carry_everything_the_metric_names():
    """Round 2: the ONE line also carries the sustained figure with its clock samples, full-callback evals/s at every N (collectives
    in the timed region), the strong-scaling legs, configs 3-5 and a parity block against the oracle -- at N = 1, 2 and 8."""
    for name, n in (("r02_bench_line.json", 1), ("r02_bench_line_2gpu.json", 2), ("r02_bench_line_4gpu.json", 4), ("r02_bench_line_8gpu.json", 8)):
        d = _load(name)
        for k in BASE + ("roofline", "clocks", "sustained", "full_callback", "strong", "configs", "parity"):
            assert k in d, (name, k)
        assert d["n_gpus"] == n and d["scaling"] == "weak" and "WEAK" in d["config"]["sharding"] or n == 1
        assert abs(d["value"] - n * 89_999_985 / (d["ms_per_step"] * 1e-3)) <= 0.02 * d["value"]
        r = d["roofline"]
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.5 < r["frac"] < 1.0 and "module" in r and "traffic_source" in r
        s = d["sustained"]
        assert s["window_s"] >= 0.5 and 0.5 < s["frac"] < 1.0 and d["clocks"]["samples"] >= 3 and "sustained leg" in d["clocks"]["window"]
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        f = d["full_callback"]
        assert f["evals_per_s"] > 1000 and "exb_eval" in f["api"] and f["separate_callbacks"]["evals_per_s"] < f["evals_per_s"]
        if n > 1:
            assert f["collectives_per_eval"] >= 1 and "replicate" in f and "owner" in f["mode"]
        st = d["strong"]
        assert st["nnzh"] == 89_999_985 and st["hess"]["ms"] > 0 and st["full_callback"]["evals_per_s"] > 0
        assert {"config4_acopf_10k", "config5_32x1e6"} <= set(d["configs"]) and (n > 1 or "config3_rocket_1e6" in d["configs"])
        p = d["parity"]
        assert p["ok"] is True and p["tolerance"] == 1e-10
        for k, v in p.items():
            if isinstance(v, dict):
                assert all(e <= 1e-10 for e in v.values() if isinstance(e, float)), (name, k)
        e = d["e2e"]
        assert 0 < e["value"] < d["value"] and e["d2h_bytes_per_step"] >= 8 * 89_999_985 * n
    one = _load("r02_bench_line.json")
    c = one["cpu_baseline"]
    assert c["kind"] == "port-compiled" and c["compiled_equals_interpreter"] is True and c["interpreter"]["value"] < c["value"]
    assert "N=10000000" in c["sample"]                       # the same workload as the reference arm
    ec = one["e2e"]["compressed"]
    assert ec["d2h_bytes_per_step"] == 8 * ec["unique_nnz"] and ec["raw_nnz_equivalent_per_s"] > 2 * one["e2e"]["value"]
    ref = _load("r02_bench_reference_arm.json")
    assert ref["impl"] == "reference" and ref["cpu_baseline"]["kind"] == "port-compiled" and ref["config"]["sample_n"] == 10_000_000
    # scaling seen in the committed lines: strong LV hess >= 6x and full evaluation >= 4x at 8 GPUs
    eight = _load("r02_bench_line_8gpu.json")
    assert one["strong"]["hess"]["ms"] / eight["strong"]["hess"]["ms"] >= 6.0
    assert one["strong"]["full_callback"]["ms_per_eval"] / eight["strong"]["full_callback"]["ms_per_eval"] >= 3.5   # 3.6-4.7 across runs
    assert one["configs"]["config5_32x1e6"]["full_callback"]["ms_per_eval"] / eight["configs"]["config5_32x1e6"]["full_callback"]["ms_per_eval"] >= 3.5   # 3.6-4.7 across runs
    assert eight["configs"]["config4_acopf_10k"]["full_callback"]["ms_per_eval"] > one["configs"]["config4_acopf_10k"]["full_callback"]["ms_per_eval"]   # negative scaling, reported


def test_traffic_file_is_tied_to_a_module_hash():
    t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    for k in ("exb_hess_g0", "exb_hessc_g0", "exb_eval_g0"):
        assert t[k]["module"].startswith("exb_") and t[k]["bytes"] > 0 and "capture" in t[k]
    assert 0.9 < t["exb_hess_g0"]["bytes"] / 879_999_864 < 1.0


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
