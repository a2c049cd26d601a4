"""Development aid: rocket nh=1e6 hess_coord! / full evaluation timing + parity of a small instance (knobs via environment)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import examodels_jl_b200 as E
from examodels_jl_b200 import models as M
from oracle.oracle_api import Oracle
core = M.goddard_rocket(1_000_000)
m = E.ExaModel(core); meta = core.meta()
x = torch.from_numpy(meta["x0"] + 0.01 * np.random.default_rng(0).uniform(-1, 1, m.nvar)).cuda()
y = torch.from_numpy(np.random.default_rng(1).standard_normal(m.ncon)).cuda()
h, j, g, c, od = m.new(m.nnzh), m.new(m.nnzj), m.new(m.nvar), m.new(m.ncon), m.new(1)
def timeit(f, n=30):
    for _ in range(5): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
th = timeit(lambda: m.hess_coord(x, y, h)); te = timeit(lambda: m.eval_all(x, y, od, g, c, j, h))
small = M.goddard_rocket(2000); o, ms = Oracle.from_core(small), E.ExaModel(small)
xs = small.meta()["x0"] + 0.01 * np.random.default_rng(0).uniform(-1, 1, o.nvar); ys = np.random.default_rng(1).standard_normal(o.ncon)
err = float(np.abs(ms.hess_coord(torch.from_numpy(xs).cuda(), torch.from_numpy(ys).cuda(), ms.new(ms.nnzh)).cpu().numpy() - o.hess_coord(xs, ys, 1.0)).max() / np.abs(o.hess_coord(xs, ys, 1.0)).max())
print(f"rocket hess {th:.4f} ms ({8 * (m.nnzh + m.nvar + m.ncon) / th / 1e6:.0f} GB/s alg)  eval {te:.4f} ms  choice {m.kernel_choice('hess')}  parity {err:.1e}")
