"""Development aid: fused evaluation (exb_eval) vs the five separate callbacks, device time."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import examodels_jl_b200 as E
from examodels_jl_b200 import models as M
key = sys.argv[1] if len(sys.argv) > 1 else "lv"
core = {"lv": lambda: M.luksan_vlcek(10_000_000), "rocket": lambda: M.goddard_rocket(1_000_000),
        "opf": lambda: M.ac_power(M.synthetic_power_data()), "family": lambda: M.pattern_family(1_000_000, 32)}[key]()
m = E.ExaModel(core)
meta = core.meta()
x = torch.from_numpy(meta["x0"] + 0.01 * np.random.default_rng(0).uniform(-1, 1, m.nvar)).cuda()
y = torch.from_numpy(np.random.default_rng(1).standard_normal(m.ncon)).cuda()
h, j, g, c, od = m.new(m.nnzh), m.new(m.nnzj), m.new(m.nvar), m.new(m.ncon), m.new(1)
def timeit(f, n=30):
    for _ in range(5): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
def sep():
    m.obj_async(x, od); m.grad(x, g); m.cons_nln(x, c); m.jac_coord(x, j); m.hess_coord(x, y, h)
t_sep = timeit(sep)
ref = [t.clone() for t in (od, g, c, j, h)]
t_fused = timeit(lambda: m.eval_all(x, y, od, g, c, j, h))
eq = [bool(torch.equal(a, b)) for a, b in zip(ref, (od, g, c, j, h))]
md = [float((a - b).abs().max() / max(1e-300, float(b.abs().max()))) for a, b in zip((od, g, c, j, h), ref)]
t_first_sep = timeit(lambda: (m.obj_async(x, od), m.grad(x, g), m.cons_nln(x, c), m.jac_coord(x, j)))
t_first = timeit(lambda: m.eval_all(x, None, od, g, c, j, None, mask=15))
t_val_sep = timeit(lambda: (m.obj_async(x, od), m.cons_nln(x, c)))
t_val = timeit(lambda: m.eval_all(x, None, od, None, c, None, None, mask=5))
print(f"{key}: first-order (obj+grad+cons+jac) separate {t_first_sep:.4f} ms  fused {t_first:.4f} ms;  values (obj+cons) separate {t_val_sep:.4f} ms  fused {t_val:.4f} ms")
alg = 8 * (m.nnzh + m.nnzj + m.ncon + m.nvar + 2 * m.nvar + m.ncon)
print(f"{key}: separate {t_sep:.4f} ms ({1e3 / t_sep:.0f} evals/s)  fused {t_fused:.4f} ms ({1e3 / t_fused:.0f} evals/s), launches {m.stats()['last_launches']}; "
      f"bitwise equal obj/grad/cons/jac/hess {eq}, max rel diff {['%.1e' % v for v in md]}")
