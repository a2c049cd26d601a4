"""Duplicate-free Hessian emitted directly by the column-tile kernel (exb_hessc_g0; the CompressedNLPModel role,
/root/reference/src/utils.jl:425-579 | ext:1290-1319) against the reference semantics restated over the oracle's raw COO:
stable sort by (col, row), unique runs, duplicates summed in slot order."""
import os

import numpy as np
import pytest

from test_gpu_products import _compress_ref
from util import assert_close, inputs

pytestmark = pytest.mark.gpu


def _models():
    from examodels_jl_b200 import models as M
    from edge_models import EDGE
    return {
        "lv_1003": lambda: M.luksan_vlcek(1003),                    # 4 tiles, ragged tail
        "lv_382": lambda: M.luksan_vlcek(382),                      # exactly one tile (128 x 3 columns - halo 2)
        "lv_5": lambda: M.luksan_vlcek(5),                          # tiny: every column is a boundary column
        "lv_guide_700": lambda: M.luksan_vlcek(700, order="guide"), # objective first
        "lv_param_300": lambda: M.luksan_vlcek_param(300),          # parameters in the objective
        "parametric": lambda: M.parametric(200),
        "only_objective": EDGE["only_objective"],
        "only_constraints": EDGE["only_constraints"],
    }


def _ref(core, sigma, with_y=True):
    from oracle.oracle_api import Oracle
    ora = Oracle.from_core(core)
    x, y = inputs(core)
    hr, hc = ora.hess_structure()
    return ora, x, y, _compress_ref(hr, hc, ora.hess_coord(x, y if with_y else None, sigma))


@pytest.mark.parametrize("name", list(_models().keys()))
def test_fused_duplicate_free_hessian(exa, name):
    import torch
    core = _models()[name]()
    ora, x, y, (rh, ch, vh) = _ref(core, 0.5)
    m = exa.ExaModel(core)
    cm = m.compressed()
    assert cm.fused_hess and exa.Plan(core).tile_info() == {"fused": True, "nnzh_unique": len(rh), "distances": exa.Plan(core).tile_info()["distances"],
                                                            "halo": exa.Plan(core).tile_info()["halo"]}
    assert cm.nnzh == len(rh) and (cm.hess_lo, cm.hess_hi) == (0, len(rh))
    for dt in (torch.int64, torch.int32):                                   # structure: closed form of the pattern shifts, `==`
        r, c = cm.new(cm.nnzh, dt).fill_(-7), cm.new(cm.nnzh, dt).fill_(-7)
        cm.hess_structure(r, c)
        assert np.array_equal(r.cpu().numpy(), rh) and np.array_equal(c.cpu().numpy(), ch)
    dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    l0 = m.stats()["launches"]
    got = cm.hess_coord(dx, dy, cm.new(cm.nnzh).fill_(float("nan")), obj_weight=0.5)
    assert_close(got.cpu().numpy(), vh, "fused duplicate-free hess")
    got2 = cm.hess_coord(dx, dy, cm.new(cm.nnzh).fill_(float("nan")), obj_weight=0.5)
    assert m.stats()["last_launches"] == 1 and torch.equal(got, got2)       # ONE launch; no atomics: bitwise reproducible
    # objective-only form (y = NULL, src/nlp.jl:1906-1915): constraint entries are structural zeros
    _, _, _, (_, _, v0) = _ref(core, 2.0, with_y=False)
    assert_close(cm.hess_coord(dx, None, cm.new(cm.nnzh).fill_(float("nan")), obj_weight=2.0).cpu().numpy(), v0, "fused, objective only")
    # host-buffer form: D2H of the unique entries only
    out = np.full(cm.nnzh, np.nan)
    m.host_hess_compressed(x, y, out, obj_weight=0.5)
    assert_close(out, vh, "exb_host_hess_compressed")
    assert m.host_bytes()[1] == 8 * cm.nnzh
    assert l0 >= 0


def test_fused_equals_the_sorted_gather_bitwise(exa):
    """The general fallback (raw COO, then a segmented sum through the (col, row)-sorted list; EXB_NO_TILE=1 disables the fused
    kernel) sums the duplicates in the same order: the two forms agree to the last bit."""
    import torch
    from examodels_jl_b200 import models as M
    core = M.luksan_vlcek(2000)
    x, y = inputs(core)
    dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    a = exa.ExaModel(core).compressed()
    os.environ["EXB_NO_TILE"] = "1"
    try:
        b = exa.ExaModel(core).compressed()
    finally:
        del os.environ["EXB_NO_TILE"]
    assert a.fused_hess and not b.fused_hess and a.nnzh == b.nnzh
    ra, ca, rb, cb = (a.new(a.nnzh, torch.int64) for _ in range(4))
    a.hess_structure(ra, ca); b.hess_structure(rb, cb)
    assert torch.equal(ra, rb) and torch.equal(ca, cb)
    va = a.hess_coord(dx, dy, a.new(a.nnzh), obj_weight=0.7)
    vb = b.hess_coord(dx, dy, b.new(b.nnzh), obj_weight=0.7)
    assert torch.equal(va, vb)


@pytest.mark.parametrize("world", [2, 3])
def test_fused_duplicate_free_hessian_on_sharded_handles(exa, world):
    """A rank owns a contiguous range of columns and evaluates whichever points touch it, so duplicates that straddle two
    shards need no exchange: the ranks' ranges tile the unique list exactly once."""
    import torch
    from examodels_jl_b200 import models as M
    core = M.luksan_vlcek(1500)
    ora, x, y, (rh, ch, vh) = _ref(core, 0.5)
    dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    cover = np.zeros(len(rh), dtype=np.int64)
    full = np.full(len(rh), np.nan)
    for r in range(world):
        m = exa.ExaModel(core, rank=r, world=world)
        cm = m.compressed()
        assert cm.fused_hess and cm.nnzh == len(rh)
        v = cm.hess_coord(dx, dy, cm.new(cm.nnzh).fill_(float("nan")), obj_weight=0.5).cpu().numpy()
        lo, hi = cm.hess_lo, cm.hess_hi
        assert not np.isnan(v[lo:hi]).any() and np.isnan(v[:lo]).all() and np.isnan(v[hi:]).all()
        cover[lo:hi] += 1
        full[lo:hi] = v[lo:hi]
        vlo, vhi = m.owned()
        assert np.all((ch[lo:hi] > vlo) & (ch[lo:hi] <= vhi))            # exactly the entries of the owned columns
    assert (cover == 1).all()
    assert_close(full, vh, "sharded fused hess")


@pytest.mark.parametrize("seed", list(range(6)) + [99])
def test_tile_kernels_on_random_shift_indexed_models(exa, seed):
    """Random patterns over different sub-ranges with shifts in [-3, 3]: the closed-form structure, the fused duplicate-free
    values (or the sorted-gather fallback when a distance's columns have a gap), the gradient and the fused evaluation against
    the oracle."""
    import torch
    from edge_models import tile_fuzz
    core = tile_fuzz(seed)
    ora, x, y, (rh, ch, vh) = _ref(core, 0.5)
    m = exa.ExaModel(core)
    cm = m.compressed()
    assert cm.nnzh == len(rh) and cm.fused_hess == exa.Plan(core).tile_info()["fused"] and cm.fused_hess == (seed != 99)
    r, c = cm.new(cm.nnzh, torch.int64), cm.new(cm.nnzh, torch.int64)
    cm.hess_structure(r, c)
    assert np.array_equal(r.cpu().numpy(), rh) and np.array_equal(c.cpu().numpy(), ch)
    dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    assert_close(cm.hess_coord(dx, dy, cm.new(cm.nnzh).fill_(float("nan")), obj_weight=0.5).cpu().numpy(), vh, "duplicate-free hess")
    assert_close(m.grad(dx, m.new(m.nvar).fill_(float("nan"))).cpu().numpy(), ora.grad(x), "grad")
    outs = [m.new(k).fill_(float("nan")) for k in (1, m.nvar, m.ncon, m.nnzj, m.nnzh)]
    m.eval_all(dx, dy, *outs, obj_weight=0.5)
    assert_close(outs[1].cpu().numpy(), ora.grad(x), "eval grad")
    assert_close(outs[4].cpu().numpy(), ora.hess_coord(x, y, 0.5), "eval hess")
    if m.ncon:
        assert_close(outs[2].cpu().numpy(), ora.cons(x), "eval cons")
        assert_close(outs[3].cpu().numpy(), ora.jac_coord(x), "eval jac")
    # sharded: owned columns tile the unique list
    if cm.fused_hess:
        full = np.full(len(rh), np.nan)
        for rk in range(3):
            ms = exa.ExaModel(core, rank=rk, world=3)
            cs = ms.compressed()
            v = cs.hess_coord(dx, dy, cs.new(cs.nnzh).fill_(float("nan")), obj_weight=0.5).cpu().numpy()
            assert np.isnan(full[cs.hess_lo:cs.hess_hi]).all()
            full[cs.hess_lo:cs.hess_hi] = v[cs.hess_lo:cs.hess_hi]
        assert_close(full, vh, "sharded duplicate-free hess")
