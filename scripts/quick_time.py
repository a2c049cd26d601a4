"""Quick device timing of every callback on LV N (development aid, not the bench)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import examodels_jl_b200 as E
from examodels_jl_b200 import models as M

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
t0 = time.time(); core = M.luksan_vlcek(N); m = E.ExaModel(core); print("create", time.time() - t0, m.stats())
meta = core.meta()
x = torch.from_numpy(meta["x0"] + 0.01 * np.random.default_rng(0).uniform(-1, 1, m.nvar)).cuda()
y = torch.from_numpy(np.random.default_rng(1).standard_normal(m.ncon)).cuda()
h, j, g, c, od = m.new(m.nnzh), m.new(m.nnzj), m.new(m.nvar), m.new(m.ncon), m.new(1)
def timeit(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return np.median(ts), np.min(ts)
for name, f, bytes_ in [
    ("hess", lambda: m.hess_coord(x, y, h), 8 * (m.nnzh + m.nvar + m.ncon)),
    ("jac", lambda: m.jac_coord(x, j), 8 * (m.nnzj + m.nvar)),
    ("grad", lambda: m.grad(x, g), 16 * m.nvar),
    ("cons", lambda: m.cons_nln(x, c), 8 * (m.ncon + m.nvar)),
    ("obj", lambda: m.obj_async(x, od), 8 * m.nvar)]:
    med, mn = timeit(f)
    print(f"{name}: median {med:.4f} ms min {mn:.4f} ms  -> {bytes_ / med / 1e6:.1f} GB/s algorithmic")
med, _ = timeit(lambda: m.hess_coord(x, y, h))
print("hess nnz/s", m.nnzh / (med * 1e-3), "kernel choice:", {cb: m.kernel_choice(cb) for cb in ("hess", "jac", "grad", "cons", "obj")})
# parity of this build / knob set against the oracle on a small instance
from oracle.oracle_api import Oracle
small = M.luksan_vlcek(3000); o, ms = Oracle.from_core(small), E.ExaModel(small)
xs = small.meta()["x0"] + 0.01 * np.random.default_rng(0).uniform(-1, 1, o.nvar); ys = np.random.default_rng(1).standard_normal(o.ncon)
dx, dy = torch.from_numpy(xs).cuda(), torch.from_numpy(ys).cuda()
rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
print("parity max rel err: hess %.2e jac %.2e grad %.2e cons %.2e obj %.2e" % (
    rel(ms.hess_coord(dx, dy, ms.new(ms.nnzh)).cpu().numpy(), o.hess_coord(xs, ys, 1.0)), rel(ms.jac_coord(dx, ms.new(ms.nnzj)).cpu().numpy(), o.jac_coord(xs)),
    rel(ms.grad(dx, ms.new(ms.nvar)).cpu().numpy(), o.grad(xs)), rel(ms.cons_nln(dx, ms.new(ms.ncon)).cpu().numpy(), o.cons(xs)),
    abs(ms.obj(dx) - o.obj(xs)) / abs(o.obj(xs))))
