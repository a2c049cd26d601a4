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
h, j, g, c = m.new(m.nnzh), m.new(m.nnzj), m.new(m.nvar), m.new(m.ncon)
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
    ("obj", lambda: m.obj(x), 8 * m.nvar)]:
    med, mn = timeit(f)
    print(f"{name}: median {med:.4f} ms min {mn:.4f} ms  -> {bytes_ / med / 1e6:.1f} GB/s algorithmic")
med, _ = timeit(lambda: m.hess_coord(x, y, h))
print("hess nnz/s", m.nnzh / (med * 1e-3))
