"""Development aid: matrix-free products on LV N (fused kernels vs the reference's sorted-structure scheme)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import examodels_jl_b200 as E
from examodels_jl_b200 import models as M
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
which = sys.argv[2] if len(sys.argv) > 2 else "lv"
core = {"lv": lambda: M.luksan_vlcek(N), "rocket": lambda: M.goddard_rocket(N), "family": lambda: M.pattern_family(N, 32),
        "opf": lambda: M.ac_power(M.synthetic_power_data())}[which]()
meta = core.meta()
def timeit(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
res = {}
for mode in ("fused", "sorted"):
    m = E.ExaModel(core, sorted_products=(mode == "sorted"))
    x = torch.from_numpy(meta["x0"] + 0.01 * np.random.default_rng(0).uniform(-1, 1, m.nvar)).cuda()
    y = torch.from_numpy(np.random.default_rng(1).standard_normal(m.ncon)).cuda()
    v = torch.from_numpy(np.random.default_rng(2).standard_normal(m.nvar)).cuda()
    w = torch.from_numpy(np.random.default_rng(3).standard_normal(m.ncon)).cuda()
    Jv, Jtw, Hv, h = m.new(m.ncon), m.new(m.nvar), m.new(m.nvar), m.new(m.nnzh)
    m.hess_coord(x, y, h); m.jac_coord(x, m.new(m.nnzj))     # tune the value kernels first (the products borrow their variant)
    res[mode] = {"jprod": timeit(lambda: m.jprod_nln(x, v, Jv)), "jtprod": timeit(lambda: m.jtprod_nln(x, w, Jtw)),
                 "hprod": timeit(lambda: m.hprod(x, y, v, Hv))}
    res[mode + "_vals"] = (Jv.clone(), Jtw.clone(), Hv.clone())
    print(which, N, mode, {k: round(t, 4) for k, t in res[mode].items()}, "ms; device MB", m.stats()["device_bytes"] / 1e6, flush=True)
    del m
for a, b, nm in zip(res["fused_vals"], res["sorted_vals"], ("jprod", "jtprod", "hprod")):
    print(nm, "fused vs sorted max rel diff", float((a - b).abs().max() / b.abs().max()))
