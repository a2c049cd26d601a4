"""Development aid: does the owner-computed gradient kernel (compute-bound, 22 registers) overlap with an HBM-bound sweep when
the two are issued on different streams?  LV N=1e7: hess_coord! + grad! back to back on one stream vs on two streams."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import examodels_jl_b200 as E
from examodels_jl_b200 import models as M
core = M.luksan_vlcek(10_000_000)
m = E.ExaModel(core)
meta = core.meta()
x = torch.from_numpy(meta["x0"] + 0.01 * np.random.default_rng(0).uniform(-1, 1, m.nvar)).cuda()
y = torch.from_numpy(np.random.default_rng(1).standard_normal(m.ncon)).cuda()
h, j, g, c, od = m.new(m.nnzh), m.new(m.nnzj), m.new(m.nvar), m.new(m.ncon), m.new(1)
m.hess_coord(x, y, h); m.grad(x, g); m.eval_all(x, y, od, g, c, j, h); torch.cuda.synchronize()
side = torch.cuda.Stream()
def timeit(f, n=30):
    for _ in range(5): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
def serial():
    m.hess_coord(x, y, h); m.grad(x, g)
def forked(first_side):
    main = torch.cuda.current_stream()
    side.wait_stream(main)
    if first_side:
        with torch.cuda.stream(side): m.grad(x, g)
        m.hess_coord(x, y, h)
    else:
        m.hess_coord(x, y, h)
        with torch.cuda.stream(side): m.grad(x, g)
    main.wait_stream(side)
print(f"hess {timeit(lambda: m.hess_coord(x, y, h)):.4f}  grad {timeit(lambda: m.grad(x, g)):.4f}  serial {timeit(serial):.4f}  "
      f"forked(grad first) {timeit(lambda: forked(True)):.4f}  forked(hess first) {timeit(lambda: forked(False)):.4f} ms")
