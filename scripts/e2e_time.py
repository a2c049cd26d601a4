"""Development aid: end-to-end (host buffers) timing of hess_coord! / jac_coord! on LV N through exb_host_*;
EXB_HOST_WINDOWS=1 disables the windowed two-stream path."""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import examodels_jl_b200 as E
from examodels_jl_b200 import models as M
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
core = M.luksan_vlcek(N); m = E.ExaModel(core)
pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
x = pin(core.meta()["x0"] + 0.01 * np.random.default_rng(0).uniform(-1, 1, m.nvar))
y = pin(np.random.default_rng(1).standard_normal(m.ncon))
h, j = pin(np.zeros(m.nnzh)), pin(np.zeros(m.nnzj))
for name, f, nnz in (("hess", lambda: m.hess_coord(x, y, h), m.nnzh), ("jac", lambda: m.jac_coord(x, j), m.nnzj)):
    f(); f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    print(f"windows={os.environ.get('EXB_HOST_WINDOWS', 'default')} {name}: median {np.median(ts) * 1e3:.3f} ms min {np.min(ts) * 1e3:.3f} ms "
          f"-> {nnz / np.median(ts):.4g} nnz/s e2e", flush=True)
