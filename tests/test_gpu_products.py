"""GPU parity of the §8f rows: matrix-free products (jprod_nln! / jtprod_nln! / hprod!) and the duplicate-free
COO (CompressedNLPModel) against the oracle."""
import numpy as np
import pytest

from util import assert_close, inputs

pytestmark = pytest.mark.gpu


def _models():
    from examodels_jl_b200 import models as M
    from edge_models import EDGE
    return {
        "lv100": lambda: M.luksan_vlcek(100),
        "lv_aug_20x3": lambda: M.luksan_vlcek_aug(20, 3),
        "opf_300": lambda: M.ac_power(M.synthetic_power_data(300, 420, 70, seed=2)),
        "rocket_50": lambda: M.goddard_rocket(50),
        "family_1000": lambda: M.pattern_family(1000, 32),
        # index expressions that differ symbolically but coincide for some points (a self loop is a DIAGONAL entry: no mirror
        # term in hprod), fixed indices shared by every point, strided / data-indexed variables
        "self_loops": EDGE["self_loops"], "mixed_gradient": EDGE["mixed_gradient"], "single_points_and_constants": EDGE["single_points_and_constants"],
        "field_types": EDGE["field_types"],
    }


@pytest.mark.parametrize("mode", ["fused", "sorted"])
@pytest.mark.parametrize("name", list(_models().keys()))
def test_products_match_oracle(exa, name, mode):
    """`fused` (default): products formed inside the derivative sweep (jprod deterministic, jtprod / hprod with FP64
    atomics); `sorted`: the reference's COO + sorted-structure SpMV (EXB_FLAG_SORTED_PRODUCTS), bitwise reproducible."""
    import torch
    from oracle.oracle_api import Oracle
    core = _models()[name]()
    ora, m = Oracle.from_core(core), exa.ExaModel(core, sorted_products=(mode == "sorted"))
    x, y = inputs(core)
    rng = np.random.default_rng(11)
    v, w = rng.standard_normal(m.nvar), rng.standard_normal(m.ncon)
    dx, dy, dv, dw = (torch.from_numpy(a).cuda() for a in (x, y, v, w))
    nan = float("nan")
    assert_close(m.jprod_nln(dx, dv, m.new(m.ncon).fill_(nan)).cpu().numpy(), ora.jprod(x, v), "jprod")
    assert_close(m.jtprod_nln(dx, dw, m.new(m.nvar).fill_(nan)).cpu().numpy(), ora.jtprod(x, w), "jtprod")
    assert_close(m.hprod(dx, dy, dv, m.new(m.nvar).fill_(nan), obj_weight=0.7).cpu().numpy(), ora.hprod(x, y, v, 0.7), "hprod")
    assert_close(m.hprod(dx, None, dv, m.new(m.nvar).fill_(nan), obj_weight=2.0).cpu().numpy(), ora.hprod(x, None, v, 2.0), "hprod obj-only")
    a = m.hprod(dx, dy, dv, m.new(m.nvar), obj_weight=0.7)
    b = m.hprod(dx, dy, dv, m.new(m.nvar), obj_weight=0.7)
    if mode == "sorted":      # sorted segmented sums, no atomics: bitwise reproducible
        assert torch.equal(a, b)
    else:                     # atomic adds: reproducible to rounding; jprod stays bitwise reproducible
        assert_close(a.cpu().numpy(), b.cpu().numpy(), "hprod twice", rtol=1e-13)
        j1 = m.jprod_nln(dx, dv, m.new(m.ncon)); j2 = m.jprod_nln(dx, dv, m.new(m.ncon))
        assert torch.equal(j1, j2)


def test_fused_products_on_a_sharded_handle(exa):
    """Partial products of the shards add up to the full product (the sorted scheme is not available on shards)."""
    import torch
    from examodels_jl_b200 import models as M
    from oracle.oracle_api import Oracle
    core = M.luksan_vlcek_aug(21, 3)
    ora = Oracle.from_core(core)
    x, y = inputs(core)
    rng = np.random.default_rng(12)
    v, w = rng.standard_normal(ora.nvar), rng.standard_normal(ora.ncon)
    dx, dy, dv, dw = (torch.from_numpy(a).cuda() for a in (x, y, v, w))
    Jv, Jtw, Hv = (torch.zeros(n, dtype=torch.float64, device="cuda") for n in (ora.ncon, ora.nvar, ora.nvar))
    for r in range(3):
        m = exa.ExaModel(core, rank=r, world=3)
        Jv += m.jprod_nln(dx, dv, m.new(m.ncon)); Jtw += m.jtprod_nln(dx, dw, m.new(m.nvar))
        Hv += m.hprod(dx, dy, dv, m.new(m.nvar), obj_weight=0.7)
    assert_close(Jv.cpu().numpy(), ora.jprod(x, v), "sharded jprod")
    assert_close(Jtw.cpu().numpy(), ora.jtprod(x, w), "sharded jtprod")
    assert_close(Hv.cpu().numpy(), ora.hprod(x, y, v, 0.7), "sharded hprod")


def test_sorted_products_on_a_sharded_handle_are_deterministic(exa):
    """EXB_FLAG_SORTED_PRODUCTS on sharded handles: every rank sorts the slots of its own points, its SpMV has a fixed summation
    order (no atomics), the partial products add up to the full product and repeat bit for bit."""
    import torch
    from examodels_jl_b200 import models as M
    from oracle.oracle_api import Oracle
    for core in (M.luksan_vlcek_aug(21, 3), M.ac_power(M.synthetic_power_data(300, 420, 70, seed=2))):
        ora = Oracle.from_core(core)
        x, y = inputs(core)
        rng = np.random.default_rng(12)
        v, w = rng.standard_normal(ora.nvar), rng.standard_normal(ora.ncon)
        dx, dy, dv, dw = (torch.from_numpy(a).cuda() for a in (x, y, v, w))
        Jv, Jtw, Hv = (torch.zeros(n, dtype=torch.float64, device="cuda") for n in (ora.ncon, ora.nvar, ora.nvar))
        for r in range(3):
            m = exa.ExaModel(core, rank=r, world=3, sorted_products=True)
            a = m.hprod(dx, dy, dv, m.new(m.nvar), obj_weight=0.7)
            b = m.hprod(dx, dy, dv, m.new(m.nvar), obj_weight=0.7)
            assert torch.equal(a, b)
            t1 = m.jtprod_nln(dx, dw, m.new(m.nvar)); t2 = m.jtprod_nln(dx, dw, m.new(m.nvar))
            assert torch.equal(t1, t2)
            Jv += m.jprod_nln(dx, dv, m.new(m.ncon)); Jtw += t1; Hv += a
        assert_close(Jv.cpu().numpy(), ora.jprod(x, v), "sharded sorted jprod")
        assert_close(Jtw.cpu().numpy(), ora.jtprod(x, w), "sharded sorted jtprod")
        assert_close(Hv.cpu().numpy(), ora.hprod(x, y, v, 0.7), "sharded sorted hprod")


def _compress_ref(rows, cols, vals):
    """src/utils.jl:478-487,564-579: stable sort of ((col,row), k) by (col,row); unique runs; sum in slot order."""
    order = np.lexsort((rows, cols))          # primary key col, secondary row; lexsort is stable
    r, c, v = rows[order], cols[order], vals[order]
    new = np.ones(len(r), dtype=bool)
    new[1:] = (r[1:] != r[:-1]) | (c[1:] != c[:-1])
    idx = np.cumsum(new) - 1
    out = np.zeros(int(new.sum()))
    for k in range(len(v)):                   # sequential, ascending slot order within a run
        out[idx[k]] += v[k]
    return r[new], c[new], out


@pytest.mark.parametrize("name", list(_models().keys()))
def test_compressed_matches_reference_semantics(exa, name):
    import torch
    from oracle.oracle_api import Oracle
    core = _models()[name]()
    ora, m = Oracle.from_core(core), exa.ExaModel(core)
    cm = m.compressed()
    x, y = inputs(core)
    dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    jr, jc = ora.jac_structure()
    hr, hc = ora.hess_structure()
    rj, cj, vj = _compress_ref(jr, jc, ora.jac_coord(x))
    rh, ch, vh = _compress_ref(hr, hc, ora.hess_coord(x, y, 0.5))
    assert (cm.nnzj, cm.nnzh) == (len(rj), len(rh))
    assert cm.nnzh <= m.nnzh and cm.nnzj <= m.nnzj
    r, c = cm.new(cm.nnzj, torch.int64), cm.new(cm.nnzj, torch.int64)
    cm.jac_structure(r, c)
    assert np.array_equal(r.cpu().numpy(), rj) and np.array_equal(c.cpu().numpy(), cj)
    r, c = cm.new(cm.nnzh, torch.int64), cm.new(cm.nnzh, torch.int64)
    cm.hess_structure(r, c)
    assert np.array_equal(r.cpu().numpy(), rh) and np.array_equal(c.cpu().numpy(), ch)
    assert_close(cm.jac_coord(dx, cm.new(cm.nnzj).fill_(float("nan"))).cpu().numpy(), vj, "compressed jac")
    assert_close(cm.hess_coord(dx, dy, cm.new(cm.nnzh).fill_(float("nan")), obj_weight=0.5).cpu().numpy(), vh, "compressed hess")


@pytest.mark.parametrize("n", [20000, 3000])
@pytest.mark.parametrize("mode", ["fused", "sorted"])
def test_targets_shared_by_thousands_of_slots(exa, n, mode):
    """A variable at a fixed index shared by every point and rows collecting thousands of augmentation terms (edge_models.
    shared_targets): the sorted (target, slot) lists have runs of n slots -- warp-summed above 64, chunked over blocks from 8192
    (n = 20000: three chunks per run; n = 3000: the warp form only) -- in grad!, cons!, the sorted products and the sorted
    duplicate-free forms; the fused products aggregate the atomics of such targets per block.  All against the oracle."""
    import torch
    from edge_models import shared_targets
    from oracle.oracle_api import Oracle
    core = shared_targets(n)
    ora, m = Oracle.from_core(core), exa.ExaModel(core, sorted_products=(mode == "sorted"))
    x, y = inputs(core)
    rng = np.random.default_rng(14)
    v, w = rng.standard_normal(m.nvar), rng.standard_normal(m.ncon)
    dx, dy, dv, dw = (torch.from_numpy(a).cuda() for a in (x, y, v, w))
    nan = float("nan")
    g1 = m.grad(dx, m.new(m.nvar).fill_(nan)); c1 = m.cons_nln(dx, m.new(m.ncon).fill_(nan))
    assert_close(g1.cpu().numpy(), ora.grad(x), "grad")
    assert_close(c1.cpu().numpy(), ora.cons(x), "cons")
    assert torch.equal(g1, m.grad(dx, m.new(m.nvar).fill_(nan))) and torch.equal(c1, m.cons_nln(dx, m.new(m.ncon).fill_(nan)))   # fixed order
    assert_close(m.jprod_nln(dx, dv, m.new(m.ncon).fill_(nan)).cpu().numpy(), ora.jprod(x, v), "jprod")
    assert_close(m.jtprod_nln(dx, dw, m.new(m.nvar).fill_(nan)).cpu().numpy(), ora.jtprod(x, w), "jtprod")
    h1 = m.hprod(dx, dy, dv, m.new(m.nvar).fill_(nan), obj_weight=0.7)
    assert_close(h1.cpu().numpy(), ora.hprod(x, y, v, 0.7), "hprod")
    assert_close(m.hprod(dx, None, dv, m.new(m.nvar).fill_(nan), obj_weight=2.0).cpu().numpy(), ora.hprod(x, None, v, 2.0), "hprod obj-only")
    if mode == "sorted":
        assert torch.equal(h1, m.hprod(dx, dy, dv, m.new(m.nvar).fill_(nan), obj_weight=0.7))
        assert torch.equal(m.jtprod_nln(dx, dw, m.new(m.nvar)), m.jtprod_nln(dx, dw, m.new(m.nvar)))
        # all five callbacks from one sweep share the same finishing kernels
        od, g, cv, jv, hv = m.new(1), m.new(m.nvar).fill_(nan), m.new(m.ncon).fill_(nan), m.new(m.nnzj), m.new(m.nnzh)
        m.eval_all(dx, dy, od, g, cv, jv, hv)
        assert torch.equal(g, g1) and torch.equal(cv, c1)
        # duplicate-free forms through the sorted list (the step variable's column: one run per (col, row) pair, short; the
        # augmentation rows' Jacobian entries are the long ones)
        cm = m.compressed()
        jr, jc = ora.jac_structure(); hr, hc = ora.hess_structure()
        rj, cj, vj = _compress_ref(jr, jc, ora.jac_coord(x))
        rh, ch, vh = _compress_ref(hr, hc, ora.hess_coord(x, y, 0.7))
        assert cm.nnzj == len(rj) and cm.nnzh == len(rh)
        assert_close(cm.jac_coord(dx, cm.new(cm.nnzj).fill_(nan)).cpu().numpy(), vj, "compressed jac")
        assert_close(cm.hess_coord(dx, dy, cm.new(cm.nnzh).fill_(nan), obj_weight=0.7).cpu().numpy(), vh, "compressed hess")
