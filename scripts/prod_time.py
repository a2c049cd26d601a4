"""Development aid: matrix-free products (jprod / jtprod / hprod), device time, fused-sweep form vs the sorted (deterministic) form."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import examodels_jl_b200 as E
from examodels_jl_b200 import models as M
key = sys.argv[1] if len(sys.argv) > 1 else "lv"
mk = {"lv": lambda: M.luksan_vlcek(10_000_000), "rocket": lambda: M.goddard_rocket(1_000_000),
      "opf": lambda: M.ac_power(M.synthetic_power_data()), "family": lambda: M.pattern_family(1_000_000, 32)}[key]
def timeit(f, n=20):
    for _ in range(4): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for kw in ({}, {"sorted_products": True}):
    core = mk()
    m = E.ExaModel(core, **kw)
    meta = core.meta()
    x = torch.from_numpy(meta["x0"] + 0.01 * np.random.default_rng(0).uniform(-1, 1, m.nvar)).cuda()
    y = torch.from_numpy(np.random.default_rng(1).standard_normal(m.ncon)).cuda()
    v = torch.from_numpy(np.random.default_rng(2).standard_normal(m.nvar)).cuda()
    jv, jtv, hv = m.new(m.ncon), m.new(m.nvar), m.new(m.nvar)
    t1 = timeit(lambda: m.jprod_nln(x, v, jv)); t2 = timeit(lambda: m.jtprod_nln(x, y, jtv)); t3 = timeit(lambda: m.hprod(x, y, v, hv))
    print(f"{key} {kw}: jprod {t1:.4f} ms  jtprod {t2:.4f} ms  hprod {t3:.4f} ms   (nvar {m.nvar}, ncon {m.ncon})", flush=True)
    del m
