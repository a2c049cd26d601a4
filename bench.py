#!/usr/bin/env python
"""bench.py -- hess_coord! throughput (nnz/s, FP64) of the B200 evaluator on Luksan-Vlcek N=10^7, plus everything
BASELINE.json's metric names next to it, in ONE JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--quick]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline (`value`, `ms_per_step`, `roofline`, `e2e`): a step = one `hess_coord!(m, x, y, hess; obj_weight=1)` over the
whole model (BASELINE.json configs[1]: LV N=10^7, nnzh = 89 999 985).  WEAK scaling: with N ranks the model is LV with
10^7 * N variables, every pattern's iterator is split into N contiguous shards, rank r evaluates shard r into its slice of
the COO buffer; there is no data-path collective (contiguous, non-overlapping slices).

Also in the line (every one measured in this run, on this box):
  sustained      the same kernel looped for >= 0.5 s (the driver's K may be a 3 ms burst), clocks sampled INSIDE that window
  full_callback  obj + grad! + cons! + jac_coord! + hess_coord! evals/s at every N, the collectives of the sharded model
                 (exb_comm_*: NCCL inside libexa_b200.so) inside the timed region; owner and replicate modes
  strong         LV N=10^7 TOTAL split over the N ranks: hess_coord! and full-callback
  configs        BASELINE configs 3 (rocket nh=10^6, N=1 only), 4 (10k-bus AC-OPF) and 5 (32 patterns x 10^6), strong-scaled:
                 hess nnz/s, roofline fraction, full-callback evals/s, max relative error against the oracle on a sample
  parity         per-rank spot check of the timed outputs against the oracle (a 2000-point window of every pattern)
  cpu_baseline   the CPU restatement on the host cores (interpreter and compiled port), N=1 only

`--impl reference` times the CPU restatement of the reference's path (oracle/, all host threads) on the same config; the
reference itself is pure Julia and cannot run in this image.
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
RTOL = 1e-10   # north star: FP64 derivative values within 1e-10 relative


def model_inputs(core):
    """x = x0 + 0.01 u (seed 0), y ~ N(0,1) (seed 1): SURVEY.md §8d."""
    meta = core.meta()
    x = meta["x0"] + 0.01 * np.random.default_rng(0).uniform(-1.0, 1.0, meta["nvar"])
    y = np.random.default_rng(1).standard_normal(meta["ncon"])
    return np.ascontiguousarray(x), np.ascontiguousarray(y)


def lv_inputs(nvar, ncon):
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
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
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


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (interpreter of the pattern IR) and its compiled form, on the host cores
# ---------------------------------------------------------------------------------------------------------------------
def cpu_hess_rates(n_points, reps=1, compiled=True):
    """hess_coord! rate of the CPU restatement (oracle/: src/hessian.jl:681-717 + nlp.jl:1917-1940) on LV N=n_points.
    Returns a dict: interpreter and (when it builds) the compiled per-pattern port, single- and multi-threaded."""
    from examodels_jl_b200 import models as M
    import oracle.oracle_api as OA
    OA.build()
    core = M.luksan_vlcek(n_points)
    ora = OA.Oracle.from_core(core)
    threads = OA.Oracle.max_threads()
    x, y = lv_inputs(ora.nvar, ora.ncon)
    out = np.zeros(ora.nnzh)

    def rate(f, thr):
        ora.set_threads(thr)
        f(x, y, 1.0, out)  # warm-up (page faults)
        best = float("inf")
        for _ in range(reps):
            t = time.perf_counter()
            f(x, y, 1.0, out)
            best = min(best, time.perf_counter() - t)
        return ora.nnzh / best, best
    res = {"n": n_points, "nnzh": ora.nnzh, "threads": threads}
    res["interp_1"], res["interp_1_s"] = rate(ora.hess_coord, 1)
    res["interp_N"], res["interp_N_s"] = rate(ora.hess_coord, threads)
    if compiled and hasattr(ora, "compile"):
        try:
            ref = out.copy()
            ora.set_threads(threads)
            ora.hess_coord(x, y, 1.0, ref)
            comp = ora.compile()
            res["compiled_1"], res["compiled_1_s"] = rate(comp.hess_coord, 1)
            res["compiled_N"], res["compiled_N_s"] = rate(comp.hess_coord, threads)
            res["compiled_equals_interpreter"] = bool(np.array_equal(out, ref))
            res["compiled_max_rel_diff"] = float(np.max(np.abs(out - ref)) / max(1e-300, np.max(np.abs(ref))))
        except Exception as e:  # the compiled arm is an extra: report why it is missing
            res["compiled_error"] = repr(e)[:200]
    return res


def cpu_baseline_entry(r):
    """`cpu_baseline` object from cpu_hess_rates: the compiled port is the stand-in for Julia's type-specialised native code
    (src/hessian.jl:681-712 is a compiled @simd loop); the interpreter is reported next to it."""
    kind = "port-compiled" if "compiled_N" in r else "port"
    val = r.get("compiled_N", r["interp_N"])
    sample = (f"LV N={r['n']} hess_coord! ({r['nnzh']} nnz), oracle/exa_oracle.cpp; "
              + (f"compiled per-pattern port (g++ -O3, emitted by the oracle from its own expanded tree): {r['threads']} threads "
                 f"{r['compiled_N_s']:.3f} s, 1 thread {r['compiled_1_s']:.3f} s; " if "compiled_N" in r else "")
              + f"IR interpreter: {r['threads']} threads {r['interp_N_s']:.3f} s, 1 thread {r['interp_1_s']:.3f} s")
    e = {"value": val, "unit": UNIT, "cores": r["threads"], "kind": kind, "sample": sample,
         "single_thread_value": r.get("compiled_1", r["interp_1"]),
         "interpreter": {"value": r["interp_N"], "single_thread_value": r["interp_1"], "kind": "port"}}
    for k in ("compiled_equals_interpreter", "compiled_max_rel_diff", "compiled_error"):
        if k in r:
            e[k] = r[k]
    return e


def run_reference(args, emit):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle.oracle_api as OA
    from examodels_jl_b200 import models as M
    OA.build()
    threads = OA.Oracle.max_threads()
    # the same workload as the GPU arm's N=1 (and its cpu_baseline): LV N=10^7, unless K + W steps of it would not end within
    # ~2.5 minutes on this box -- then a bounded sample of the same two patterns
    cal = cpu_hess_rates(200_000, compiled=True)
    rate = cal.get("compiled_N", cal["interp_N"])
    budget_s = 150.0 / (args.steps + args.warmup)
    n = int(min(N_PER_GPU, max(20_000, rate / 9.0 * budget_s)))
    core = M.luksan_vlcek(n)
    ora = OA.Oracle.from_core(core)
    ora.set_threads(threads)
    x, y = lv_inputs(ora.nvar, ora.ncon)
    out = np.zeros(ora.nnzh)
    kind, f = "port", ora.hess_coord
    if "compiled_N" in cal:
        kind, f = "port-compiled", ora.compile().hess_coord
    steps, warmup = args.steps, args.warmup
    for _ in range(warmup):
        f(x, y, 1.0, out)
    t0 = time.perf_counter()
    for _ in range(steps):
        f(x, y, 1.0, out)
    dt = (time.perf_counter() - t0) / steps
    val = ora.nnzh / dt
    sample = (f"LV N={n} ({ora.nnzh} nnz) per step" + (" = the full configs[1] workload" if n == N_PER_GPU else ", same patterns as N=10^7")
              + f"; oracle/exa_oracle.cpp ({kind}), {threads} host threads")
    emit({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "Luksan-Vlcek N=10^7 hess_coord! (configs[1])", "sample_n": n,
                   "note": ("the GPU arm at N ranks evaluates N x 10^7 points (weak scaling); this arm always evaluates one "
                            "10^7-point model on rank 0: a rate is compared with a rate")},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample,
                         "interpreter_value": cal["interp_N"], "interpreter_single_thread": cal["interp_1"]},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
class Ctx:
    pass


def timed(ctx, fn, steps, warmup=3):
    """`steps` calls of fn on torch's current stream between two CUDA events, barrier + synchronize on both sides;
    returns (ms per step, max over ranks; this rank's ms per step)."""
    torch = ctx.torch
    for _ in range(warmup):
        fn()
    ctx.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.current_stream()
    a.record(st)
    for _ in range(steps):
        fn()
    b.record(st)
    ctx.barrier()
    mine = a.elapsed_time(b) / steps
    return ctx.maxr(mine), mine


def alg_bytes_hess(m, shards, core):
    """8 nnzh (each output word once) + 8 nvar (x once) + 8 ncon (y once) + iterator bytes, for the part of the
    model this rank evaluates."""
    nnz = sum(s["hess_hi"] - s["hess_lo"] for s in shards)
    itr = 0
    rows = 0
    for p, s in zip(core.patterns, shards):
        n = s["hi"] - s["lo"]
        if p.itr.range is None:
            itr += n * p.itr.array.dtype.itemsize
        if p.kind == 1:
            rows += n
    frac = max((s["hi"] - s["lo"]) / max(1, p.nitr) for p, s in zip(core.patterns, shards)) if shards else 1.0
    return int(8 * nnz + 8 * m.nvar * min(1.0, frac) + 8 * rows + itr)


def spot_check(ctx, m, core, x_h, y_h, outs, npts=2000, sigma=1.0):
    """Per-rank parity of what the timed calls wrote: the first `npts` LOCAL points of every pattern against the oracle run on
    the windowed model (nlp.window_core: slots depend on their own point only, so the windowed model's outputs are slices of
    the full model's).  `outs`: dict with any of hess / jac / cons device tensors.  Returns max relative error
    (norm-wise per vector, the rule of tests/util.assert_close)."""
    import examodels_jl_b200 as E
    from examodels_jl_b200.nlp import window_core, window_slices
    from oracle.oracle_api import Oracle
    P = E.Plan(core)
    fi = [P.pattern_info(k) for k in range(P.npatterns())]
    wins = []
    for k in range(len(fi)):
        s = m.shard(k)
        wins.append((s["lo"], min(s["hi"], s["lo"] + npts)))
    wc = window_core(core, wins)
    Pw = E.Plan(wc)
    wi = [Pw.pattern_info(k) for k in range(Pw.npatterns())]
    sl = window_slices(fi, wi, wins)
    ora = Oracle.from_core(wc)
    yw = np.zeros(ora.ncon)
    for d in sl:
        if "rows" in d:
            yw[d["rows"][1]] = y_h[d["rows"][0]]
    res = {}
    refs = {"hess": lambda: ora.hess_coord(x_h, yw, sigma), "jac": lambda: ora.jac_coord(x_h), "cons": lambda: ora.cons(x_h)}
    for name, t in outs.items():
        ref = refs[name]()
        got = np.empty_like(ref)
        key = "rows" if name == "cons" else name
        for d in sl:
            if key in d and d["n"] > 0:
                a, b = d[key]
                got[b] = t[a.start:a.stop].cpu().numpy()
        scale = float(np.max(np.abs(ref))) if ref.size else 0.0
        err = float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), scale))) if ref.size and scale > 0 else 0.0
        res[name] = err
    res["points_per_pattern"] = npts
    return res


def full_check(m, core, x_h, y_h, outs, sigma=1.0):
    """Whole-vector parity against the oracle (small models: AC-OPF)."""
    from oracle.oracle_api import Oracle
    ora = Oracle.from_core(core)
    refs = {"hess": lambda: ora.hess_coord(x_h, y_h, sigma), "jac": lambda: ora.jac_coord(x_h), "cons": lambda: ora.cons(x_h),
            "grad": lambda: ora.grad(x_h), "obj": lambda: np.array([ora.obj(x_h)])}
    res = {}
    for name, t in outs.items():
        ref = refs[name]()
        got = t.cpu().numpy() if hasattr(t, "cpu") else np.asarray(t)
        scale = float(np.max(np.abs(ref))) if ref.size else 0.0
        res[name] = float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), scale))) if ref.size and scale > 0 else 0.0
    return res


def eval_legs(ctx, core, name, steps=20, check="window", graph=False, peak=None):
    """One model, sharded over the job's ranks: hess_coord! (no collective) and the full five-callback evaluation with the
    collectives of the sharded model inside the timed region (owner mode: sharded consumer; replicate mode: every vector
    complete on every rank), plus parity of what was written against the oracle."""
    torch, E = ctx.torch, ctx.E
    t0 = time.time()
    m = E.ExaModel(core, device=ctx.local, rank=ctx.rank, world=ctx.world)
    build_s = time.time() - t0
    if ctx.world > 1:
        m.comm_init(mode="owner")
    x_h, y_h = model_inputs(core)
    x, y = torch.from_numpy(x_h).cuda(), torch.from_numpy(y_h).cuda()
    hess, jac, c, g, od = m.new(m.nnzh), m.new(m.nnzj), m.new(m.ncon), m.new(m.nvar), m.new(1)
    shards = [m.shard(k) for k in range(m.npatterns)]

    def sepcb():   # the five callbacks one by one, back to back on one stream (the reference's protocol, runbenchmark.jl:79-101)
        m.obj_async(x, od); m.grad(x, g); m.cons_nln(x, c); m.jac_coord(x, jac); m.hess_coord(x, y, hess)

    def allcb():   # the same five outputs from ONE sweep (exb_eval: every data point evaluated once)
        m.eval_all(x, y, od, g, c, jac, hess)
    sepcb(); allcb()   # first calls tune (and synchronise)
    torch.cuda.synchronize()
    ms_h, _ = timed(ctx, lambda: m.hess_coord(x, y, hess), steps)
    ms_s, _ = timed(ctx, sepcb, steps)
    l0, c0 = m.stats()["launches"], m.comm_stats()["collectives"]
    ms_f, _ = timed(ctx, allcb, steps)
    nl = (m.stats()["launches"] - l0) // (steps + 3)
    ncoll = (m.comm_stats()["collectives"] - c0) // (steps + 3)
    bi = m.build_info()
    out = {"model": name, "nvar": m.nvar, "ncon": m.ncon, "nnzj": m.nnzj, "nnzh": m.nnzh, "build_s": round(build_s, 2),
           "build": {"nvcc_s": round(bi["nvcc_s"], 2), "module_cached": bool(m.stats()["module_cached"]), "load_upload_sort_s": round(bi["load_s"], 3),
                     "tune_s": round(bi["tune_s"], 3)},
           "hess": {"ms": ms_h, "nnz_per_s": m.nnzh / (ms_h * 1e-3), "collective": "none (contiguous disjoint slices per rank)"},
           "full_callback": {"ms_per_eval": ms_f, "evals_per_s": 1e3 / ms_f, "launches_per_eval": int(nl),
                             "api": "exb_eval(EXB_EVAL_ALL): one sweep", "callbacks": "obj+grad!+cons!+jac_coord!+hess_coord!",
                             "separate_callbacks_ms_per_eval": ms_s, "separate_callbacks_evals_per_s": 1e3 / ms_s}}
    # the first-order evaluation a solver does at a new iterate (obj + grad! + cons! + jac_coord!) from one sweep
    ms_1, _ = timed(ctx, lambda: m.eval_all(x, None, od, g, c, jac, None, mask=15), steps)
    ms_1s, _ = timed(ctx, lambda: (m.obj_async(x, od), m.grad(x, g), m.cons_nln(x, c), m.jac_coord(x, jac)), steps)
    out["first_order"] = {"ms_per_eval": ms_1, "evals_per_s": 1e3 / ms_1, "api": "exb_eval(EXB_EVAL_FIRST): one sweep",
                          "callbacks": "obj+grad!+cons!+jac_coord!", "separate_callbacks_ms_per_eval": ms_1s}
    if hasattr(m, "compressed"):
        try:
            cm = m.compressed()
            if cm.fused_hess:
                vc = cm.new(cm.nnzh)
                cm.hess_coord(x, y, vc)
                ms_c, _ = timed(ctx, lambda: cm.hess_coord(x, y, vc), steps)
                out["hess_duplicate_free"] = {"ms": ms_c, "unique_nnz": cm.nnzh, "raw_nnz_equivalent_per_s": m.nnzh / (ms_c * 1e-3),
                                              "unique_nnz_per_s": cm.nnzh / (ms_c * 1e-3), "api": "exb_hess_compressed: one launch (column-tile kernel)"}
                del vc
        except Exception as ex:
            out["hess_duplicate_free"] = {"error": repr(ex)[:200]}
    ab = alg_bytes_hess(m, shards, core)
    if peak:
        out["hess"]["algorithmic_bytes_per_rank"] = ab
        out["hess"]["frac_of_hbm_peak"] = ab / (ms_h * 1e-3) / 1e9 / peak
    if ctx.world > 1:
        out["full_callback"]["mode"] = "owner (sharded consumer: g on owned variables, c / jac / hess on own points)"
        out["full_callback"]["collectives_per_eval"] = int(ncoll)
        m.comm_set_mode("replicate")
        allcb()
        c0 = m.comm_stats()["collectives"]
        ms_r, _ = timed(ctx, allcb, steps)
        out["full_callback_replicate"] = {"ms_per_eval": ms_r, "evals_per_s": 1e3 / ms_r,
                                          "collectives_per_eval": int((m.comm_stats()["collectives"] - c0) // (steps + 3)),
                                          "mode": "replicate (obj, g, c complete on every rank; jac / hess stay sharded)"}
    if graph and ctx.world == 1:
        gr = m.capture_full_eval(x, y, od, g, c, jac, hess)
        ms_g, _ = timed(ctx, gr.replay, steps)
        out["full_callback"]["cuda_graph_ms_per_eval"] = ms_g
        out["full_callback"]["cuda_graph_evals_per_s"] = 1e3 / ms_g
        out["full_callback"]["cuda_graph_what"] = "the five separate callbacks on parallel graph branches"
        gf = m.capture_fused_eval(x, y, od, g, c, jac, hess)
        ms_gf, _ = timed(ctx, gf.replay, steps)
        out["full_callback"]["cuda_graph_fused_ms_per_eval"] = ms_gf
        out["full_callback"]["cuda_graph_fused_evals_per_s"] = 1e3 / ms_gf
    # parity of what the timed calls left in the buffers
    allcb()
    torch.cuda.synchronize()
    if check == "window":
        out["parity"] = spot_check(ctx, m, core, x_h, y_h, {"hess": hess, "jac": jac, "cons": c})
        out["parity"]["against"] = "oracle on the windowed model (first 2000 local points of every pattern, this rank)"
    elif check == "full":   # replicate mode (set above when world > 1): g, c, obj are complete
        outs = {"grad": g, "cons": c, "obj": od}
        if ctx.world > 1:
            m.gather_coo(1, jac); m.gather_coo(2, hess)
        outs.update({"hess": hess, "jac": jac})
        torch.cuda.synchronize()
        out["parity"] = full_check(m, core, x_h, y_h, outs)
        out["parity"]["against"] = "oracle, whole vectors, all five callbacks"
    errs = [v for v in out.get("parity", {}).values() if isinstance(v, float)]
    worst = ctx.maxr(max(errs) if errs else 0.0)
    out["parity"]["max_rel_err_over_ranks"] = worst
    out["parity"]["ok"] = bool(worst <= RTOL)
    if ctx.world > 1:
        m.comm_destroy()
    del m
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)   # ~0.3 s timed region: long enough for nvidia-smi samples inside it
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--n", type=int, default=N_PER_GPU, help="points per GPU (default: the BASELINE config)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--quick", action="store_true", help="headline only: skip the strong / configs legs")
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
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    E.build_library()

    ctx = Ctx()
    ctx.torch, ctx.E, ctx.rank, ctx.world, ctx.local = torch, E, rank, world, local

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    ctx.barrier, ctx.maxr = barrier, maxr

    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak, peak_src = float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        pass

    # ---------------- headline: LV N = 10^7 per GPU, hess_coord! (weak scaling, no collective) ----------------
    n_total = args.n * world
    core = M.luksan_vlcek(n_total)
    m = E.ExaModel(core, device=local, rank=rank, world=world)
    if world > 1:
        m.comm_init(mode="owner")
    xh, yh = lv_inputs(m.nvar, m.ncon)
    xp, yp = torch.from_numpy(xh).pin_memory(), torch.from_numpy(yh).pin_memory()
    x, y = xp.cuda(non_blocking=True), yp.cuda(non_blocking=True)
    hess = m.new(m.nnzh)
    shards = [m.shard(k) for k in range(m.npatterns)]
    alg_bytes = alg_bytes_hess(m, shards, core)

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
    total_ms = maxr(ev[0].elapsed_time(ev[-1]))
    per = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    ms_per_step = total_ms / args.steps
    value = m.nnzh / (ms_per_step * 1e-3)
    kernel_ms = float(np.mean(per))  # one generated kernel per step: the launch duration incl. launch gap

    # sustained: the same kernel for >= 0.5 s, so that the clock samples lie INSIDE a timed window whatever K the caller chose
    sus_steps = max(200, int(0.6 / (ms_per_step * 1e-3)))
    barrier()
    s_wall0 = time.time()
    sus_ms, sus_mine = timed(ctx, lambda: m.hess_coord(x, y, hess, obj_weight=1.0), sus_steps, warmup=0)
    s_wall1 = time.time()
    sustained = {"steps": sus_steps, "ms_per_step": sus_ms, "value": m.nnzh / (sus_ms * 1e-3), "unit": UNIT,
                 "frac": alg_bytes / (sus_mine * 1e-3) / 1e9 / peak, "window_s": round(s_wall1 - s_wall0, 3)}
    clocks = None
    if rank == 0:
        time.sleep(0.12)   # let the sampler's last lines arrive
        timed_clk = sampler.summary(t_wall0, t_wall1)
        clocks = sampler.summary(s_wall0, s_wall1) or {}
        clocks["window"] = f"sustained leg ({sus_steps} steps, {s_wall1 - s_wall0:.2f} s), sampled every 50 ms inside it"
        clocks["timed_region_s"] = round(t_wall1 - t_wall0, 4)
        clocks["timed_region_samples"] = timed_clk["samples"] if timed_clk else 0
        if timed_clk:
            clocks["timed_region_sm_mhz"] = timed_clk["sm_mhz"]

    # parity of the timed output: per-rank window against the oracle
    parity = {"lv_weak": spot_check(ctx, m, core, xh, yh, {"hess": hess})}

    # full-callback evals/s at every N on the headline model, collectives inside the timed region
    g, c, j, od = m.new(m.nvar), m.new(m.ncon), m.new(m.nnzj), m.new(1)

    def sepcb():
        m.obj_async(x, od); m.grad(x, g); m.cons_nln(x, c); m.jac_coord(x, j); m.hess_coord(x, y, hess)

    def allcb():
        m.eval_all(x, y, od, g, c, j, hess)
    sepcb(); allcb()
    torch.cuda.synchronize()
    mss, _ = timed(ctx, sepcb, 20)
    c0, l1 = m.comm_stats()["collectives"], m.stats()["launches"]
    msf, _ = timed(ctx, allcb, 20)
    full = {"evals_per_s": 1e3 / msf, "ms_per_eval": msf, "callbacks": "obj+grad!+cons!+jac_coord!+hess_coord!",
            "api": "exb_eval(EXB_EVAL_ALL): one sweep, every data point evaluated once",
            "separate_callbacks": {"evals_per_s": 1e3 / mss, "ms_per_eval": mss, "note": "the five C-ABI callbacks back to back on one stream"},
            "model": f"LV N={n_total} ({args.n} per GPU, weak)", "launches_per_eval": (m.stats()["launches"] - l1) // 23}
    if world > 1:
        full["mode"] = "owner (sharded consumer)"
        full["collectives_per_eval"] = (m.comm_stats()["collectives"] - c0) // 23
        full["collectives"] = "obj: ncclAllReduce of 8 bytes; grad!: none (owner-computed per variable); cons!/jac/hess: none (own rows / slices)"
        m.comm_set_mode("replicate")
        allcb()
        msr, _ = timed(ctx, allcb, 10)
        full["replicate"] = {"evals_per_s": 1e3 / msr, "ms_per_eval": msr,
                             "collectives": "obj all-reduce (8 B) + all-gather of g (nvar doubles) + all-gather of c (ncon doubles); jac / hess stay sharded"}
        m.comm_set_mode("owner")
    allcb()
    torch.cuda.synchronize()
    pw = spot_check(ctx, m, core, xh, yh, {"jac": j, "cons": c})
    parity["lv_weak"].update({k: v for k, v in pw.items() if k in ("jac", "cons")})

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
    e2e_dt = maxr((time.perf_counter() - t0) / e2e_steps)
    e2e_val = m.nnzh / e2e_dt
    hb = torch.tensor(m.host_bytes(), dtype=torch.float64, device="cuda")   # bytes the library itself copied in the last call
    if world > 1:
        dist.all_reduce(hb, op=dist.ReduceOp.SUM)
    h2d_bytes, d2h_bytes = int(hb[0].item()), int(hb[1].item())
    ok = bool(np.isfinite(hh[shards[0]["hess_lo"]:shards[0]["hess_hi"]]).all())
    e2e = {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
           "api": "exb_host_hess (C ABI, pinned host buffers)", "steps": e2e_steps, "finite": ok, "ms_per_step": e2e_dt * 1e3}
    # the same through the duplicate-free form (the CompressedNLPModel role): D2H of the unique entries only
    if world == 1 and hasattr(m, "host_hess_compressed"):
        try:
            cm = m.compressed()
            hc = torch.empty(cm.nnzh, dtype=torch.float64).pin_memory().numpy()
            m.host_hess_compressed(xn, yn, hc)
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                m.host_hess_compressed(xn, yn, hc)
            dtc = (time.perf_counter() - t0) / e2e_steps
            b = m.host_bytes()
            e2e["compressed"] = {"api": "exb_host_hess_compressed (unique lower-triangle coordinates, duplicates summed on the device)",
                                 "ms_per_step": dtc * 1e3, "unique_nnz": cm.nnzh, "unique_nnz_per_s": cm.nnzh / dtc,
                                 "raw_nnz_equivalent_per_s": m.nnzh / dtc, "h2d_bytes_per_step": b[0], "d2h_bytes_per_step": b[1]}
        except Exception as ex:
            e2e["compressed"] = {"error": repr(ex)[:200]}

    choice = m.kernel_choice("hess")   # the first-call tuner's verdict: launch-shape variant, classic or persistent form
    hess_kernel = "exb_hessp_g0" if choice["persistent"] else "exb_hess_g0"
    module = os.path.basename(E.Plan(M.luksan_vlcek(1000)).module_path())
    if world > 1:
        m.comm_destroy()
    del m, hess, g, c, j, x, y
    torch.cuda.empty_cache()

    # ---------------- strong scaling and the other BASELINE configs ----------------
    strong, configs = None, None
    if not args.quick:
        strong = eval_legs(ctx, M.luksan_vlcek(N_PER_GPU), f"LV N={N_PER_GPU} total over {world} rank(s)", graph=True, peak=peak)
        configs = {}
        configs["config4_acopf_10k"] = eval_legs(ctx, M.ac_power(M.synthetic_power_data()), "synthetic 10k-bus AC-OPF pattern set (15 patterns)",
                                                 check="full", graph=True, peak=peak)
        configs["config4_acopf_10k"]["note"] = ("latency-bound (0.64 M nnz): sharding it makes it slower -- every collective (obj all-reduce, "
                                                "grad / cons all-reduce over nvar / ncon because indices are iterator data) costs more than the kernels")
        configs["config5_32x1e6"] = eval_legs(ctx, M.pattern_family(1_000_000, 32), "32 distinct patterns x 10^6 points (AoS iterators)", peak=peak)
        if world == 1:
            configs["config3_rocket_1e6"] = eval_legs(ctx, M.goddard_rocket(1_000_000), "COPS Goddard rocket nh=10^6 (formulation parity-unpinned: not in the reference tree)",
                                                      peak=peak)
        for k_, v_ in list(configs.items()) + [("lv_strong", strong)]:
            parity[k_] = v_.pop("parity")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # traffic: a STATIC figure from the committed ncu capture, tied to the hash of the module it was taken on
    traffic, traffic_note = None, "no capture on file"
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f).get(hess_kernel)
        if isinstance(tj, dict):
            if tj.get("module") == module:
                traffic, traffic_note = tj["bytes"], f"static: ncu --set full capture {tj.get('capture')} of this very module ({module})"
            else:
                traffic_note = f"STALE capture ignored: taken on module {tj.get('module')}, running {module}"
                print("bench.py: profiles/traffic.json is stale for " + hess_kernel + ": " + traffic_note, file=sys.stderr)
    except Exception:
        pass
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9

    cpu = None
    if world == 1 and not args.no_cpu:
        cpu = cpu_baseline_entry(cpu_hess_rates(N_PER_GPU if args.n == N_PER_GPU else min(args.n, N_PER_GPU), reps=1))

    parity["tolerance"] = RTOL
    parity["ok"] = all(v.get("ok", all(e <= RTOL for e in v.values() if isinstance(e, float))) for v in parity.values() if isinstance(v, dict))
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"Luksan-Vlcek N={args.n} per GPU (BASELINE configs[1]), hess_coord! only; nvar={n_total} nnzh={9 * n_total - 15}",
                   "sharding": (f"WEAK: {world} contiguous iterator shards of a {world}x larger model, no collective in hess_coord!; "
                                "`strong` / `configs` in this line are STRONG (fixed total size)") if world > 1 else "single GPU",
                   "cpu_binding": (f"rank 0 bound to {len(near)} cores local to its GPU (NVML affinity)" if near else "none"),
                   "l2": f"per step {alg_bytes / 1e6:.0f} MB of inputs+outputs > L2 ({L2_BYTES / 1e6:.0f} MB); no explicit flush",
                   "inputs": "x = x0 + 0.01 U(-1,1) seed 0; y ~ N(0,1) seed 1; obj_weight = 1"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_note, "kernel": hess_kernel, "module": module, "kernel_choice": choice,
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kernel_ms, "peak_source": peak_src,
                     "sustained_frac": sustained["frac"]},
        "sustained": sustained,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "full_callback": full,
        "parity": parity,
    }
    if strong:
        out["strong"] = strong
    if configs:
        out["configs"] = configs
    if cpu:
        out["cpu_baseline"] = cpu
    emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
