"""Development aid: A/B timing of hess_coord! on LV N=1e7 under different EXB_TUNE_* environments, interleaved."""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import examodels_jl_b200 as E
from examodels_jl_b200 import models as M
N = 10_000_000
envs = [dict(kv.split("=") for kv in a.split(",") if kv) for a in sys.argv[1:]]
core = M.luksan_vlcek(N)
x = torch.from_numpy(core.meta()["x0"] + 0.01 * np.random.default_rng(0).uniform(-1, 1, N)).cuda()
y = torch.from_numpy(np.random.default_rng(1).standard_normal(N - 2)).cuda()
models = []
for e in envs:
    for k in list(os.environ):
        if k.startswith("EXB_TUNE_"): del os.environ[k]
    os.environ.update(e)
    models.append(E.ExaModel(core))
h = models[0].new(models[0].nnzh)
for rnd in range(4):
    for e, m in zip(envs, models):
        for _ in range(20): m.hess_coord(x, y, h)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(300): m.hess_coord(x, y, h)
        b.record(); torch.cuda.synchronize()
        print(rnd, e, f"{a.elapsed_time(b) / 300:.4f} ms", flush=True)
