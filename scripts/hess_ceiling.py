"""Development aid: the headline kernel's memory ceiling -- the LV slot structure (6 + 3 slots per point, same bytes) with a
body that has no transcendentals, against LV itself."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import examodels_jl_b200 as E
from examodels_jl_b200 import models as M
from examodels_jl_b200.nlp import ExaCore
N = 10_000_000
def cheap(N):
    c = ExaCore(); x = c.add_var(N, start=M.lv_x0(N))
    c.add_con(lambda i: x[i] ** 2 * x[i + 1] ** 2 * x[i + 2] ** 2, range(1, N - 1))
    c.add_obj(lambda i: 100 * (x[i - 1] ** 2 - x[i]) ** 2 + (x[i - 1] - 1) ** 2, range(2, N + 1))
    return c
for name, core in (("cheap body, LV slot structure", cheap(N)), ("LV", M.luksan_vlcek(N))):
    m = E.ExaModel(core); meta = core.meta()
    x = torch.from_numpy(meta["x0"] + 0.01 * np.random.default_rng(0).uniform(-1, 1, m.nvar)).cuda()
    y = torch.from_numpy(np.random.default_rng(1).standard_normal(m.ncon)).cuda()
    h = m.new(m.nnzh)
    for _ in range(10): m.hess_coord(x, y, h)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50): m.hess_coord(x, y, h)
    b.record(); torch.cuda.synchronize()
    t = a.elapsed_time(b) / 50
    print(f"{name}: nnzh {m.nnzh}  hess {t:.4f} ms -> {8 * (m.nnzh + m.nvar + m.ncon) / t / 1e6:.0f} GB/s algorithmic, choice {m.kernel_choice('hess')}")
    del m, h
