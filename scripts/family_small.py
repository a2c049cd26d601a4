"""Development aid: hess_coord! of the 32-pattern family at several sizes on ONE GPU (fixed per-launch costs vs size)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import examodels_jl_b200 as E
from examodels_jl_b200 import models as M
for n in (125_000, 250_000, 500_000, 1_000_000):
    core = M.pattern_family(n, 32); m = E.ExaModel(core); meta = core.meta()
    x = torch.from_numpy(meta["x0"] + 0.01 * np.random.default_rng(0).uniform(-1, 1, m.nvar)).cuda()
    y = torch.from_numpy(np.random.default_rng(1).standard_normal(m.ncon)).cuda()
    h = m.new(m.nnzh)
    for _ in range(5): m.hess_coord(x, y, h)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50): m.hess_coord(x, y, h)
    b.record(); torch.cuda.synchronize()
    t = a.elapsed_time(b) / 50
    print(f"n={n}: hess {t:.4f} ms  -> {t / n * 1e9:.1f} fs/point/pattern-set, choice {m.kernel_choice('hess')}")
    del m, h
