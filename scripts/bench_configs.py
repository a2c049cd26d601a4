"""Measures every callback on the BASELINE.json configs 2-5 on one B200 and spot-checks parity against
the oracle on a reduced instance of the same model.  Writes gpurun_out/configs_<tag>.json + a markdown table."""
import json, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import examodels_jl_b200 as E
from examodels_jl_b200 import models as M
from oracle.oracle_api import Oracle

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
which = sys.argv[2].split(",") if len(sys.argv) > 2 else ["lv", "rocket", "opf", "family"]


def timeit(f, n=30):
    for _ in range(3): f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def parity(small):
    o, m = Oracle.from_core(small), E.ExaModel(small)
    meta = small.meta()
    x = meta["x0"] + 0.01 * np.random.default_rng(0).uniform(-1, 1, o.nvar)
    y = np.random.default_rng(1).standard_normal(o.ncon)
    dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    def rel(a, b):
        s = np.abs(b).max() if b.size else 0.0
        return float(np.abs(a - b).max() / s) if s > 0 else 0.0
    return {"hess": rel(m.hess_coord(dx, dy, m.new(m.nnzh)).cpu().numpy(), o.hess_coord(x, y, 1.0)),
            "jac": rel(m.jac_coord(dx, m.new(m.nnzj)).cpu().numpy(), o.jac_coord(x)),
            "grad": rel(m.grad(dx, m.new(m.nvar)).cpu().numpy(), o.grad(x)),
            "cons": rel(m.cons_nln(dx, m.new(m.ncon)).cpu().numpy(), o.cons(x)),
            "obj": abs(m.obj(dx) - o.obj(x)) / max(abs(o.obj(x)), 1e-300)}


CFG = {
    "lv": ("config 2: Luksan-Vlcek N=1e7", lambda: M.luksan_vlcek(10_000_000), lambda: M.luksan_vlcek(5000)),
    "rocket": ("config 3: Goddard rocket nh=1e6", lambda: M.goddard_rocket(1_000_000), lambda: M.goddard_rocket(2000)),
    "opf": ("config 4: synthetic AC-OPF 10k buses / 14k branches / 2.5k gens", lambda: M.ac_power(M.synthetic_power_data()),
            lambda: M.ac_power(M.synthetic_power_data(500, 700, 125))),
    "family": ("config 5: 32 patterns x 1e6 points", lambda: M.pattern_family(1_000_000, 32), lambda: M.pattern_family(2000, 32)),
}
out = []
for key in which:
    name, big, small = CFG[key]
    t0 = time.time(); core = big(); t_front = time.time() - t0
    t0 = time.time(); m = E.ExaModel(core); t_build = time.time() - t0
    meta = core.meta()
    x = torch.from_numpy(meta["x0"] + 0.01 * np.random.default_rng(0).uniform(-1, 1, m.nvar)).cuda()
    y = torch.from_numpy(np.random.default_rng(1).standard_normal(m.ncon)).cuda()
    h, j, g, c, od = m.new(m.nnzh), m.new(m.nnzj), m.new(m.nvar), m.new(m.ncon), m.new(1)
    itb = sum(b.nbytes for b in core.to_ir()[1]) if key != "lv" else 0   # iterator bytes (host AoS size)
    cbs = {"hess": (lambda: m.hess_coord(x, y, h), 8 * (m.nnzh + m.nvar + m.ncon)),
           "jac": (lambda: m.jac_coord(x, j), 8 * (m.nnzj + m.nvar)),
           "grad": (lambda: m.grad(x, g), 16 * m.nvar),
           "cons": (lambda: m.cons_nln(x, c), 8 * (m.ncon + m.nvar)),
           "obj": (lambda: m.obj_async(x, od), 8 * m.nvar)}
    res = {"config": name, "nvar": m.nvar, "ncon": m.ncon, "nnzj": m.nnzj, "nnzh": m.nnzh, "npatterns": m.npatterns,
           "front_end_s": t_front, "build_s": t_build, "iterator_bytes_host": itb}
    tot = 0.0
    for cb, (f, nbytes) in cbs.items():
        med, mn = timeit(f)
        tot += med
        res[cb] = {"ms": med, "ms_min": mn, "alg_GBps_excl_iter": nbytes / med / 1e6}
    def allcb():
        m.obj_async(x, od); m.grad(x, g); m.cons_nln(x, c); m.jac_coord(x, j); m.hess_coord(x, y, h)
    med, _ = timeit(allcb, 20)
    res["full_eval_ms"] = med; res["evals_per_s"] = 1e3 / med; res["hess_nnz_per_s"] = m.nnzh / (res["hess"]["ms"] * 1e-3)
    gr = m.capture_full_eval(x, y, od, g, c, j, h)
    med_g, _ = timeit(gr.replay, 20)
    res["full_eval_graph_ms"] = med_g; res["evals_per_s_graph"] = 1e3 / med_g
    res["launches_per_full_eval"] = None
    l0 = m.stats()["launches"]; allcb(); torch.cuda.synchronize(); res["launches_per_full_eval"] = m.stats()["launches"] - l0
    del gr, m, h, j
    torch.cuda.empty_cache()
    res["parity_max_rel_err_small_instance"] = parity(small())
    print(json.dumps(res), flush=True)
    out.append(res)
json.dump(out, open(f"gpurun_out/configs_{tag}.json", "w"), indent=1)
