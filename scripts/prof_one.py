"""Development aid: run one callback of one config a few times (for ncu)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import examodels_jl_b200 as E
from examodels_jl_b200 import models as M
key, cb = sys.argv[1], sys.argv[2]
core = {"lv": lambda: M.luksan_vlcek(10_000_000), "rocket": lambda: M.goddard_rocket(1_000_000),
        "opf": lambda: M.ac_power(M.synthetic_power_data()), "family": lambda: M.pattern_family(1_000_000, 32)}[key]()
m = E.ExaModel(core)
meta = core.meta()
x = torch.from_numpy(meta["x0"] + 0.01 * np.random.default_rng(0).uniform(-1, 1, m.nvar)).cuda()
y = torch.from_numpy(np.random.default_rng(1).standard_normal(m.ncon)).cuda()
h, j, g, c, od = m.new(m.nnzh), m.new(m.nnzj), m.new(m.nvar), m.new(m.ncon), m.new(1)
cm = m.compressed() if cb == "hessc" else None
vc = cm.new(cm.nnzh) if cm else None
f = {"eval": lambda: m.eval_all(x, y, od, g, c, j, h), "hessc": lambda: cm.hess_coord(x, y, vc), "hess": lambda: m.hess_coord(x, y, h), "jac": lambda: m.jac_coord(x, j), "grad": lambda: m.grad(x, g),
     "cons": lambda: m.cons_nln(x, c), "obj": lambda: m.obj_async(x, od)}[cb]
for _ in range(12): f()
torch.cuda.synchronize()
