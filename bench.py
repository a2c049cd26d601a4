#!/usr/bin/env python
"""bench.py -- hess_coord! throughput (nnz/s, FP64) of the B200 evaluator on Luksan-Vlcek N=10^7.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one `hess_coord!(m, x, y, hess; obj_weight=1)` over the whole model (BASELINE.json
configs[1]: LV N=10^7, nnzh = 89 999 985).  Per GPU the workload is fixed (weak scaling): with N
ranks the model is LV with 10^7 * N variables, every pattern's iterator is split into N contiguous
shards, rank r evaluates shard r into its slice of the COO buffer; there is no data-path collective
(contiguous, non-overlapping slices).  `value` = total nnz / max-over-ranks time.

`--impl reference` times the CPU restatement of the reference's path (oracle/, all host threads) on
the same config; the reference itself is pure Julia and cannot run in this image.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "sparse Lagrangian Hessian nnz/sec (FP64), hess_coord! on Luksan-Vlcek"
UNIT = "nnz/s"
N_PER_GPU = 10_000_000
L2_BYTES = 126 * 2 ** 20


def lv_inputs(nvar, ncon):
    """x = x0 + 0.01 u (seed 0), y ~ N(0,1) (seed 1): SURVEY.md §8d."""
    i = np.arange(1, nvar + 1)
    x0 = np.where(i % 2 == 1, -1.2, 1.0)
    x = x0 + 0.01 * np.random.default_rng(0).uniform(-1.0, 1.0, nvar)
    y = np.random.default_rng(1).standard_normal(ncon)
    return x, y


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the GPU is under load (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [f.strip() for f in line.split(",")]))

    def stop(self):
        if self.proc:
            self.proc.terminate()

    def summary(self, t0, t1):
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 and len(r) >= 8]
        if not rows:
            return None
        sm = sorted(float(r[1]) for r in rows)
        reasons = [n for k, n in ((4, "hw_slowdown"), (5, "hw_thermal_slowdown"), (6, "sw_thermal_slowdown"),
                                  (7, "sw_power_cap")) if any(r[k].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": reasons, "samples": len(rows),
                "power_w_max": max(float(r[3]) for r in rows)}


def cpu_hess_rate(n_points, threads, reps=1):
    """Oracle (CPU restatement of src/hessian.jl:681-717 + nlp.jl:1917-1940) hess_coord! rate on LV."""
    from examodels_jl_b200 import models as M
    from oracle.oracle_api import Oracle
    core = M.luksan_vlcek(n_points)
    ora = Oracle.from_core(core)
    ora.set_threads(threads)
    x, y = lv_inputs(ora.nvar, ora.ncon)
    out = np.zeros(ora.nnzh)
    ora.hess_coord(x, y, 1.0, out)  # warm-up (page faults)
    best = float("inf")
    for _ in range(reps):
        t = time.perf_counter()
        ora.hess_coord(x, y, 1.0, out)
        best = min(best, time.perf_counter() - t)
    return ora.nnzh / best, best


def run_reference(args, emit):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle.oracle_api as OA
    from examodels_jl_b200 import models as M
    OA.build()
    threads = OA.Oracle.max_threads()
    # bounded sample of the LV N=10^7 workload: calibrate, then size a step so that the W + K steps take ~2 minutes
    rate, _ = cpu_hess_rate(200_000, threads)
    per_step_s = min(1.0, 120.0 / (args.steps + args.warmup))
    n = int(min(N_PER_GPU, max(20_000, rate / 9.0 * per_step_s)))
    core = M.luksan_vlcek(n)
    ora = OA.Oracle.from_core(core)
    ora.set_threads(threads)
    x, y = lv_inputs(ora.nvar, ora.ncon)
    out = np.zeros(ora.nnzh)
    steps, warmup = args.steps, args.warmup
    for _ in range(warmup):
        ora.hess_coord(x, y, 1.0, out)
    t0 = time.perf_counter()
    for _ in range(steps):
        ora.hess_coord(x, y, 1.0, out)
    dt = (time.perf_counter() - t0) / steps
    val = ora.nnzh / dt
    sample = f"LV N={n} ({ora.nnzh} nnz) per step, same patterns as N=10^7; oracle/exa_oracle.cpp, {threads} host threads"
    emit({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "Luksan-Vlcek N=10^7 hess_coord! (configs[1]); CPU arm runs a bounded sample", "sample_n": n},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)   # ~0.35 s timed region: long enough for nvidia-smi samples inside it
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--n", type=int, default=N_PER_GPU, help="points per GPU (default: the BASELINE config)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    # stdout carries exactly ONE line, the JSON: anything a library prints on fd 1 meanwhile (NCCL's version banner, ...)
    # is sent to stderr, and fd 1 is restored just before the line is written
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(obj), flush=True)

    if args.impl == "reference":
        return run_reference(args, emit)

    import torch
    import torch.distributed as dist
    import examodels_jl_b200 as E
    from examodels_jl_b200 import models as M

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    near = None
    if world > 1:   # one process per GPU: stay on the cores (and memory) next to this GPU before any pinned buffer exists
        from examodels_jl_b200.parallel import bind_near_gpu
        near = bind_near_gpu(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    E.build_library()

    n_total = args.n * world
    core = M.luksan_vlcek(n_total)
    m = E.ExaModel(core, device=local, rank=rank, world=world)
    xh, yh = lv_inputs(m.nvar, m.ncon)
    xp, yp = torch.from_numpy(xh).pin_memory(), torch.from_numpy(yh).pin_memory()
    x, y = xp.cuda(non_blocking=True), yp.cuda(non_blocking=True)
    hess = m.new(m.nnzh)
    # slice of the COO buffer this rank writes (contiguous per pattern)
    shards = [m.shard(k) for k in range(m.npatterns)]
    local_nnz = sum(s["hess_hi"] - s["hess_lo"] for s in shards)
    # algorithmic bytes per launch on this rank: output words once + x once + y once
    local_pts = max(s["hi"] - s["lo"] for s in shards)
    alg_bytes = 8 * local_nnz + 8 * (local_pts + 2) + 8 * local_pts

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        import atexit
        sampler.start()
        atexit.register(sampler.stop)
    stream = torch.cuda.current_stream()
    for _ in range(args.warmup):
        m.hess_coord(x, y, hess, obj_weight=1.0)
    barrier()
    l0 = m.stats()["launches"]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    t_wall0 = time.time()
    ev[0].record(stream)
    for k in range(args.steps):
        m.hess_coord(x, y, hess, obj_weight=1.0)
        ev[k + 1].record(stream)
    barrier()
    t_wall1 = time.time()
    launches = m.stats()["launches"] - l0
    total_ms = ev[0].elapsed_time(ev[-1])
    per = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = m.nnzh / (ms_per_step * 1e-3)
    kernel_ms = float(np.mean(per))  # one generated kernel per step: the launch duration incl. launch gap

    # clocks: the timed region may be shorter than one nvidia-smi sample; if so, keep the same kernel
    # running (untimed) until enough samples exist and say so
    clocks = None
    if rank == 0:
        clocks = sampler.summary(t_wall0, t_wall1)
        probe = "timed region"
        if clocks is None or clocks["samples"] < 3:
            p0 = time.time()
            while time.time() - p0 < 1.5:
                for _ in range(200):
                    m.hess_coord(x, y, hess, obj_weight=1.0)
                torch.cuda.synchronize()
            clocks = sampler.summary(p0 + 0.3, time.time())
            probe = "same kernel looped for 1.5 s right after the timed region (timed region shorter than the sampling period)"
        sampler.stop()
        if clocks is not None:
            clocks["window"] = probe

    # e2e: the reference-facing C-ABI call with HOST buffers (exb_host_hess): H2D of x and y from pinned memory,
    # kernel, D2H of the whole hess vector -- inside the timed region, every step
    if world == 1:
        hh = torch.empty(m.nnzh, dtype=torch.float64).pin_memory().numpy()
    else:  # page-lock only the slices this rank receives
        hh = np.empty(m.nnzh)
        rt = torch.cuda.cudart()
        for s_ in shards:
            a0 = (hh.ctypes.data + 8 * s_["hess_lo"]) & ~4095
            a1 = (hh.ctypes.data + 8 * s_["hess_hi"] + 4095) & ~4095
            hh[s_["hess_lo"]:s_["hess_hi"]] = 0.0
            assert int(rt.cudaHostRegister(a0, a1 - a0, 0)) == 0
    xn, yn = xp.numpy(), yp.numpy()
    e2e_steps = max(3, min(args.steps, 10))
    m.hess_coord(xn, yn, hh, obj_weight=1.0)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        m.hess_coord(xn, yn, hh, obj_weight=1.0)
    barrier()
    e2e_dt = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_dt, op=dist.ReduceOp.MAX)
    e2e_val = m.nnzh / float(e2e_dt.item())
    hb = torch.tensor(m.host_bytes(), dtype=torch.float64, device="cuda")   # bytes the library itself copied in the last call
    if world > 1:
        dist.all_reduce(hb, op=dist.ReduceOp.SUM)
    h2d_bytes, d2h_bytes = int(hb[0].item()), int(hb[1].item())
    ok = bool(np.isfinite(hh[shards[0]["hess_lo"]:shards[0]["hess_hi"]]).all())

    # the other four callbacks, back to back on one stream (metric 2: full-callback evals/s)
    full = None
    if world == 1:
        g, c, j = m.new(m.nvar), m.new(m.ncon), m.new(m.nnzj)
        od = m.new(1)

        def allcb():
            m.obj_async(x, od); m.grad(x, g); m.cons_nln(x, c); m.jac_coord(x, j); m.hess_coord(x, y, hess)
        for _ in range(3):
            allcb()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(20):
            allcb()
        b.record(stream)
        torch.cuda.synchronize()
        msf = a.elapsed_time(b) / 20
        full = {"evals_per_s": 1e3 / msf, "ms_per_eval": msf, "callbacks": "obj+grad!+cons!+jac_coord!+hess_coord!"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_src = None, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
        peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        peak = 6650.0
    traffic = None
    choice = m.kernel_choice("hess")   # the first-call tuner's verdict: launch-shape variant, classic or persistent form
    hess_kernel = "exb_hessp_g0" if choice["persistent"] else "exb_hess_g0"
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(hess_kernel)
    except Exception:
        pass
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9

    cpu = None
    if world == 1 and not args.no_cpu:
        import oracle.oracle_api as OA
        OA.build()
        threads = OA.Oracle.max_threads()
        n_s = 1_000_000
        r1, t1 = cpu_hess_rate(n_s, 1)
        rN, tN = cpu_hess_rate(n_s, threads, reps=2)
        cpu = {"value": rN, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"LV N={n_s} hess_coord! (same two patterns as N=10^7), oracle/exa_oracle.cpp interpreting the pattern IR; "
                         f"{threads} threads {tN:.3f} s; single thread {r1:.4g} nnz/s ({t1:.3f} s)",
               "single_thread_value": r1}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"Luksan-Vlcek N={args.n} per GPU (BASELINE configs[1]), hess_coord! only; nvar={m.nvar} ncon={m.ncon} nnzh={m.nnzh}",
                   "sharding": f"{world} contiguous iterator shards, no collective" if world > 1 else "single GPU",
                   "cpu_binding": (f"rank 0 bound to {len(near)} cores local to its GPU (NVML affinity)" if near else "none"),
                   "l2": f"per step {alg_bytes / 1e6:.0f} MB of inputs+outputs > L2 ({L2_BYTES / 1e6:.0f} MB); no explicit flush",
                   "inputs": "x = x0 + 0.01 U(-1,1) seed 0; y ~ N(0,1) seed 1; obj_weight = 1"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": hess_kernel, "kernel_choice": choice, "algorithmic_bytes_per_launch": alg_bytes,
                     "kernel_ms": kernel_ms, "peak_source": peak_src},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "api": "exb_host_hess (C ABI, pinned host buffers)", "steps": e2e_steps, "finite": ok},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if cpu:
        out["cpu_baseline"] = cpu
    if full:
        out["full_callback"] = full
    emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
