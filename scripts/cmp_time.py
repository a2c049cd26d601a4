"""Development aid: device timing of the duplicate-free Hessian (fused column-tile kernel vs raw + sorted gather) on LV N."""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import examodels_jl_b200 as E
from examodels_jl_b200 import models as M
N = int(float(sys.argv[1])) if len(sys.argv) > 1 and sys.argv[1] != "family" else 10_000_000
core = M.pattern_family(1_000_000, 32) if len(sys.argv) > 1 and sys.argv[1] == "family" else M.luksan_vlcek(N)
meta = core.meta()
def timeit(f, n=30):
    for _ in range(5): f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))
x = torch.from_numpy(meta["x0"] + 0.01 * np.random.default_rng(0).uniform(-1, 1, meta["nvar"])).cuda()
y = torch.from_numpy(np.random.default_rng(1).standard_normal(meta["ncon"])).cuda()
m = E.ExaModel(core); cm = m.compressed()
v = cm.new(cm.nnzh)
med, mn = timeit(lambda: cm.hess_coord(x, y, v))
alg = 8 * (cm.nnzh + m.nvar + m.ncon)
print(f"fused={cm.fused_hess} unique nnzh {cm.nnzh} of {m.nnzh}: median {med:.4f} ms min {mn:.4f} ms -> {alg / med / 1e6:.0f} GB/s algorithmic, "
      f"{m.nnzh / med / 1e6:.1f} G raw-nnz-equivalent/s, launches {m.stats()['last_launches']}")
h = m.new(m.nnzh)
med2, mn2 = timeit(lambda: m.hess_coord(x, y, h))
print(f"raw hess_coord!: median {med2:.4f} ms min {mn2:.4f}")
if len(sys.argv) > 2:
    os.environ["EXB_NO_TILE"] = "1"
    m2 = E.ExaModel(core); c2 = m2.compressed(); v2 = c2.new(c2.nnzh)
    med3, mn3 = timeit(lambda: c2.hess_coord(x, y, v2))
    print(f"raw + sorted gather: median {med3:.4f} ms min {mn3:.4f}; equal to fused: {bool(torch.equal(v, v2))}")
