"""Development aid: time hess_coord! on LV N for several launch shapes (EXB_TUNE_* knobs)."""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import examodels_jl_b200 as E
from examodels_jl_b200 import models as M

name, N = sys.argv[1].split(":")
N = int(float(N))
cfgs = [tuple(int(v) for v in a.split(",")) for a in sys.argv[2:]] or [None]
def skel(N):
    """LV-shaped output (6 + 3 slots per point) with trivial arithmetic: the store skeleton of the hess kernel."""
    c = E.ExaCore(); x = c.add_var(N, start=M.lv_x0(N))
    c.add_con(lambda i: x[i] * x[i + 1] + x[i + 1] * x[i + 2] + x[i] * x[i + 2], range(1, N - 1))
    c.add_obj(lambda i: 100 * (x[i - 1] ** 2 - x[i]) ** 2 + (x[i - 1] - 1) ** 2, range(2, N + 1))
    return c
core = {"lv": lambda: M.luksan_vlcek(N), "skel": lambda: skel(N), "rocket": lambda: M.goddard_rocket(N), "family": lambda: M.pattern_family(N, 32),
        "opf": lambda: M.ac_power(M.synthetic_power_data(N, int(1.4 * N), N // 4))}[name]()
meta = core.meta()
x = torch.from_numpy(meta["x0"] + 0.01 * np.random.default_rng(0).uniform(-1, 1, meta["nvar"])).cuda()
y = torch.from_numpy(np.random.default_rng(1).standard_normal(meta["ncon"])).cuda()
for cfg in cfgs:
    blk, minb = cfg if cfg else ("auto", "auto")
    if cfg:
        os.environ["EXB_TUNE_BLOCK"], os.environ["EXB_TUNE_MINB"] = str(blk), str(minb)
    m = E.ExaModel(core)
    h = m.new(m.nnzh); j = m.new(m.nnzj)
    res = {}
    for name, f in (("hess", lambda: m.hess_coord(x, y, h)), ("jac", lambda: m.jac_coord(x, j))):
        for _ in range(5): f()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(100): f()
        b.record(); torch.cuda.synchronize()
        res[name] = a.elapsed_time(b) / 100
    print(f"block {blk} minb {minb}: hess {res['hess']:.4f} ms ({8 * (m.nnzh + m.nvar + m.ncon) / res['hess'] / 1e6:.0f} GB/s alg, nnzh {m.nnzh})  jac {res['jac']:.4f} ms", flush=True)
    del m
