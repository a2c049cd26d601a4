"""exb_eval: obj + grad! + cons! + jac_coord! + hess_coord! at the same x from ONE sweep (every data point evaluated once by
exb_eval_g0; the composition the reference performs callback by callback, /root/reference/src/nlp.jl:1827-1940), against the
oracle and against the separate callbacks."""
import numpy as np
import pytest

from util import assert_close, inputs

pytestmark = pytest.mark.gpu


def _models():
    from examodels_jl_b200 import models as M
    from edge_models import EDGE
    d = {
        "lv_1003": lambda: M.luksan_vlcek(1003),
        "lv_guide_300": lambda: M.luksan_vlcek(300, order="guide"),
        "lv_aug_20x3": lambda: M.luksan_vlcek_aug(20, 3),
        "opf_300": lambda: M.ac_power(M.synthetic_power_data(300, 420, 70, seed=2)),
        "rocket_50": lambda: M.goddard_rocket(50),
        "family_1000": lambda: M.pattern_family(1000, 32),
        "parametric": lambda: M.parametric(200),
    }
    d.update({"edge_" + k: f for k, f in EDGE.items()})
    return d


def _separate(m, dx, dy, w):
    nan = float("nan")
    od = m.new(1).fill_(nan)
    m.obj_async(dx, od)
    return (od, m.grad(dx, m.new(m.nvar).fill_(nan)), m.cons_nln(dx, m.new(m.ncon).fill_(nan)), m.jac_coord(dx, m.new(m.nnzj).fill_(nan)),
            m.hess_coord(dx, dy, m.new(m.nnzh).fill_(nan), obj_weight=w))


@pytest.mark.parametrize("name", list(_models().keys()))
def test_fused_evaluation_matches_oracle_and_separate_callbacks(exa, name):
    import torch
    from oracle.oracle_api import Oracle
    core = _models()[name]()
    ora, m = Oracle.from_core(core), exa.ExaModel(core)
    x, y = inputs(core)
    dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    nan = float("nan")
    outs = [m.new(n).fill_(nan) for n in (1, m.nvar, m.ncon, m.nnzj, m.nnzh)]
    m.eval_all(dx, dy, *outs, obj_weight=0.5)      # first call: the tuner tries every launch-shape variant of the sweep
    outs = [m.new(n).fill_(nan) for n in (1, m.nvar, m.ncon, m.nnzj, m.nnzh)]
    m.eval_all(dx, dy, *outs, obj_weight=0.5)
    od, g, c, j, h = outs
    ref = ora.obj(x)
    assert abs(float(od.item()) - ref) <= 1e-10 * max(1.0, abs(ref))
    assert_close(g.cpu().numpy(), ora.grad(x), "eval grad")
    assert_close(c.cpu().numpy(), ora.cons(x), "eval cons")
    assert_close(j.cpu().numpy(), ora.jac_coord(x), "eval jac")
    assert_close(h.cpu().numpy(), ora.hess_coord(x, y, 0.5), "eval hess")
    # the same words the separate callbacks write: same generated code for every slot (tight tolerance: the compiler may contract
    # a*b+c differently inside the fused body); obj: a different partition of the partial sums
    sep = _separate(m, dx, dy, 0.5)
    for a, b, what in zip(outs[1:], sep[1:], ("grad", "cons", "jac", "hess")):
        assert_close(a.cpu().numpy(), b.cpu().numpy(), "fused vs separate " + what, rtol=1e-13)
    assert abs(float(od.item()) - float(sep[0].item())) <= 1e-12 * max(1.0, abs(ref))
    # reproducible (for a fixed launch-shape variant: no atomics anywhere)
    outs2 = [m.new(n).fill_(nan) for n in (1, m.nvar, m.ncon, m.nnzj, m.nnzh)]
    m.eval_all(dx, dy, *outs2, obj_weight=0.5)
    assert all(torch.equal(a, b) for a, b in zip(outs, outs2))
    # objective-only Hessian form (y = NULL, src/nlp.jl:1906-1915) and partial masks (callbacks one by one)
    h0 = m.new(m.nnzh).fill_(nan)
    m.eval_all(dx, None, m.new(1), m.new(m.nvar), m.new(m.ncon), m.new(m.nnzj), h0, obj_weight=2.0)
    assert_close(h0.cpu().numpy(), ora.hess_coord(x, None, 2.0), "eval hess, objective only")
    c2, j2 = m.new(m.ncon).fill_(nan), m.new(m.nnzj).fill_(nan)
    m.eval_all(dx, None, None, None, c2, j2, None, mask=4 | 8)
    assert torch.equal(c2, sep[2]) and torch.equal(j2, sep[3])
    # first-order sweep (obj | grad | cons | jac: exb_eval1_g0) and value sweep (obj | cons: exb_eval0_g0)
    o1, g1, c1, j1 = m.new(1).fill_(nan), m.new(m.nvar).fill_(nan), m.new(m.ncon).fill_(nan), m.new(m.nnzj).fill_(nan)
    m.eval_all(dx, None, o1, g1, c1, j1, None, mask=15)
    m.eval_all(dx, None, o1, g1, c1, j1, None, mask=15)
    assert abs(float(o1.item()) - ref) <= 1e-10 * max(1.0, abs(ref))
    assert_close(g1.cpu().numpy(), ora.grad(x), "first-order sweep grad")
    assert_close(c1.cpu().numpy(), ora.cons(x), "first-order sweep cons")
    assert_close(j1.cpu().numpy(), ora.jac_coord(x), "first-order sweep jac")
    o0, c0 = m.new(1).fill_(nan), m.new(m.ncon).fill_(nan)
    m.eval_all(dx, None, o0, None, c0, None, None, mask=5)
    assert abs(float(o0.item()) - ref) <= 1e-10 * max(1.0, abs(ref))
    assert_close(c0.cpu().numpy(), ora.cons(x), "value sweep cons")
    assert torch.equal(c0, sep[2])          # the value function is the very code of exb_cons_g0


def test_fused_evaluation_is_one_sweep(exa):
    """LV: the fused call is 2 launches (the sweep -- which also writes g, the model having one shift-indexed objective pattern --
    and the fixed-order sum of the objective partials) against 6 for the five separate callbacks, and bitwise equal to them
    where the generated slot code is shared.  With EXB_NO_EGRAD-style models (several objective patterns) the owner-computed
    gradient kernel is the third launch."""
    import torch
    from examodels_jl_b200 import models as M
    core = M.luksan_vlcek(50_000)
    m = exa.ExaModel(core)
    x, y = inputs(core)
    dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    outs = [m.new(n) for n in (1, m.nvar, m.ncon, m.nnzj, m.nnzh)]
    m.eval_all(dx, dy, *outs)     # tunes
    m.eval_all(dx, dy, *outs)
    assert m.stats()["last_launches"] == 2
    sep = _separate(m, dx, dy, 1.0)
    same = {w: bool(torch.equal(a, b)) for a, b, w in zip(outs[1:], sep[1:], ("grad", "cons", "jac", "hess"))}
    print("bitwise equal to the separate callbacks:", same)
    assert same["grad"]           # same first-order slots summed in the same order (ascending point, then slot)


def test_fused_evaluation_on_sharded_handles(exa):
    """Shards of the fused sweep add up (no communicator: partial results, zero outside what a rank computes)."""
    import torch
    from examodels_jl_b200 import models as M
    from oracle.oracle_api import Oracle
    for core in (M.luksan_vlcek(1500), M.luksan_vlcek_aug(21, 3), M.ac_power(M.synthetic_power_data(300, 420, 70, seed=2))):
        ora = Oracle.from_core(core)
        x, y = inputs(core)
        dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
        tot = [torch.zeros(n, dtype=torch.float64, device="cuda") for n in (1, ora.nvar, ora.ncon)]
        jac = torch.full((ora.nnzj,), float("nan"), dtype=torch.float64, device="cuda")
        hess = torch.full((ora.nnzh,), float("nan"), dtype=torch.float64, device="cuda")
        for r in range(3):
            m = exa.ExaModel(core, rank=r, world=3)
            od, g, c = m.new(1), m.new(m.nvar), m.new(m.ncon)
            m.eval_all(dx, dy, od, g, c, jac, hess, obj_weight=0.5)      # jac / hess: every rank writes its own slices
            tot[0] += od; tot[1] += g; tot[2] += c
        assert abs(float(tot[0].item()) - ora.obj(x)) <= 1e-10 * max(1.0, abs(ora.obj(x)))
        assert_close(tot[1].cpu().numpy(), ora.grad(x), "sharded eval grad")
        assert_close(tot[2].cpu().numpy(), ora.cons(x), "sharded eval cons")
        assert_close(jac.cpu().numpy(), ora.jac_coord(x), "sharded eval jac")
        assert_close(hess.cpu().numpy(), ora.hess_coord(x, y, 0.5), "sharded eval hess")


def _egrad_models():
    from examodels_jl_b200 import models as M
    from edge_models import EDGE, tile_fuzz
    return {
        "lv_5": lambda: M.luksan_vlcek(5),                  # one ragged block: first and last at once
        "lv_129": lambda: M.luksan_vlcek(129),
        "lv_1003": lambda: M.luksan_vlcek(1003),
        "lv_guide_700": lambda: M.luksan_vlcek(700, order="guide"),
        "lv_100k": lambda: M.luksan_vlcek(100_003),         # interior blocks (unchecked gather) + a ragged tail
        "lv_param_300": lambda: M.luksan_vlcek_param(300),
        "only_objective": EDGE["only_objective"],
        "fuzz_99": lambda: tile_fuzz(99),                   # objective over range(100, 140) of 150+ variables: the rest of g is zero-filled
        "fuzz_0": lambda: tile_fuzz(0), "fuzz_1": lambda: tile_fuzz(1), "fuzz_3": lambda: tile_fuzz(3),
    }


@pytest.mark.parametrize("name", list(_egrad_models().keys()))
def test_gradient_written_by_the_sweep_itself(exa, name):
    """A model with ONE shift-indexed objective pattern: exb_eval (all five, and the first-order form) writes g from the
    first-order slots its sweep already holds -- no gradient launch.  Against the oracle, against grad! (same summation order:
    ascending point, then slot), and the launch count drops by one."""
    import torch
    from oracle.oracle_api import Oracle
    core = _egrad_models()[name]()
    plan = exa.Plan(core)
    nobj1 = sum(1 for k in range(plan.npatterns()) if plan.pattern_info(k)["kind"] == 0 and plan.pattern_info(k)["o1step"] > 0)
    has = "EGRAD = true" in plan.source()
    assert has == (nobj1 == 1), (nobj1, has)
    ora, m = Oracle.from_core(core), exa.ExaModel(core)
    x, y = inputs(core)
    dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    nan = float("nan")
    g_sep = m.grad(dx, m.new(m.nvar).fill_(nan))
    for mask in (31, 15):
        outs = [m.new(n).fill_(nan) for n in (1, m.nvar, m.ncon, m.nnzj, m.nnzh)]
        for _ in range(2):   # first call tunes
            outs[1].fill_(nan)
            m.eval_all(dx, dy if mask == 31 else None, outs[0], outs[1], outs[2], outs[3], outs[4] if mask == 31 else None, mask=mask)
        launches = m.stats()["last_launches"]
        assert_close(outs[1].cpu().numpy(), ora.grad(x), f"sweep gradient (mask {mask})")
        assert_close(outs[1].cpu().numpy(), g_sep.cpu().numpy(), f"sweep gradient vs grad! (mask {mask})", rtol=1e-13)
        if has:
            assert launches == 2 + (1 if m.nconaug > 0 else 0), launches      # sweep + objective sum: no gradient launch
            if name.startswith("lv"):   # same slots, same order; bitwise where the compiler contracts a*b+c alike in d012 / d01 and d1
                assert torch.equal(outs[1], g_sep)
