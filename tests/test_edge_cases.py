"""Edge cases (empty / ragged iterators, one-row patterns, constant and Null bodies, numerically coinciding
indices, mixed field types): the oracle against finite differences on CPU, the CUDA path against the oracle on GPU."""
import numpy as np
import pytest

from edge_models import EDGE
from util import assert_close, inputs


@pytest.mark.parametrize("name", list(EDGE))
def test_oracle_edge_models_vs_finite_differences(name):
    from oracle.oracle_api import Oracle
    import examodels_jl_b200 as E
    core = EDGE[name]()
    o, p = Oracle.from_core(core), E.Plan(core)
    assert (p.nvar, p.ncon, p.nnzj, p.nnzh) == (o.nvar, o.ncon, o.nnzj, o.nnzh)
    for k in range(p.npatterns()):
        assert p.pattern_info(k) == o.pattern_info(k)
    x, y = inputs(core, 2)
    h = 1e-6
    eye = np.eye(o.nvar)
    gfd = np.array([(o.obj(x + h * e) - o.obj(x - h * e)) / (2 * h) for e in eye])
    np.testing.assert_allclose(o.grad(x), gfd, rtol=1e-5, atol=1e-6)
    jr, jc = o.jac_structure()
    J = np.zeros((o.ncon, o.nvar)); np.add.at(J, (jr - 1, jc - 1), o.jac_coord(x))
    if o.ncon:
        Jfd = np.array([(o.cons(x + h * e) - o.cons(x - h * e)) / (2 * h) for e in eye]).T
        np.testing.assert_allclose(J, Jfd, rtol=1e-5, atol=1e-6)
    hr, hc = o.hess_structure()
    assert (hr >= hc).all()
    L = np.zeros((o.nvar, o.nvar)); np.add.at(L, (hr - 1, hc - 1), o.hess_coord(x, y, 0.7))
    H = L + np.tril(L, -1).T

    def glag(z):
        Jz = np.zeros((o.ncon, o.nvar)); np.add.at(Jz, (jr - 1, jc - 1), o.jac_coord(z))
        return 0.7 * o.grad(z) + Jz.T @ y
    Hfd = np.array([(glag(x + h * e) - glag(x - h * e)) / (2 * h) for e in eye])
    np.testing.assert_allclose(H, Hfd, rtol=1e-4, atol=1e-5 * max(1.0, np.abs(Hfd).max()))


def test_self_loop_slots_are_doubled():
    from oracle.oracle_api import Oracle
    core = EDGE["self_loops"]()
    o = Oracle.from_core(core)
    info = o.pattern_info(0)
    hr, hc = o.hess_structure()
    d = core.patterns[0].itr.array
    k = 0                                           # point 0 is a self loop: f == t
    assert d["f"][k] == d["t"][k]
    rows = hr[info["o2"] + k * info["o2step"]: info["o2"] + (k + 1) * info["o2step"]]
    cols = hc[info["o2"] + k * info["o2step"]: info["o2"] + (k + 1) * info["o2step"]]
    assert info["o2step"] == len(rows) and (rows == cols).sum() > 2      # symbolic dedupe keeps every slot


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(EDGE))
def test_gpu_edge_models_match_oracle(exa, name):
    import torch
    from oracle.oracle_api import Oracle
    core = EDGE[name]()
    ora, m = Oracle.from_core(core), exa.ExaModel(core)
    x, y = inputs(core, 2)
    dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    nan = float("nan")
    ref = ora.obj(x)
    assert abs(m.obj(dx) - ref) <= 1e-10 * max(1.0, abs(ref))
    assert_close(m.grad(dx, m.new(m.nvar).fill_(nan)).cpu().numpy(), ora.grad(x), "grad")
    assert_close(m.cons_nln(dx, m.new(m.ncon).fill_(nan)).cpu().numpy(), ora.cons(x), "cons")
    assert_close(m.jac_coord(dx, m.new(m.nnzj).fill_(nan)).cpu().numpy(), ora.jac_coord(x), "jac")
    assert_close(m.hess_coord(dx, dy, m.new(m.nnzh).fill_(nan), obj_weight=0.7).cpu().numpy(), ora.hess_coord(x, y, 0.7), "hess")
    jr, jc = ora.jac_structure(); hr, hc = ora.hess_structure()
    r, c = m.new(m.nnzj, torch.int64), m.new(m.nnzj, torch.int64); m.jac_structure(r, c)
    assert np.array_equal(r.cpu().numpy(), jr) and np.array_equal(c.cpu().numpy(), jc)
    r, c = m.new(m.nnzh, torch.int32), m.new(m.nnzh, torch.int32); m.hess_structure(r, c)
    assert np.array_equal(r.cpu().numpy().astype(np.int64), hr) and np.array_equal(c.cpu().numpy().astype(np.int64), hc)
    if m.ncon and m.nnzj:
        v = torch.from_numpy(np.random.default_rng(3).standard_normal(m.nvar)).cuda()
        assert_close(m.jprod_nln(dx, v, m.new(m.ncon)).cpu().numpy(), ora.jprod(x, v.cpu().numpy()), "jprod")
        assert_close(m.hprod(dx, dy, v, m.new(m.nvar), obj_weight=0.7).cpu().numpy(), ora.hprod(x, y, v.cpu().numpy(), 0.7), "hprod")
    # host-buffer shims on ragged sizes
    hh = np.full(m.nnzh, np.nan); assert_close(m.hess_coord(x, y, hh, obj_weight=0.7), ora.hess_coord(x, y, 0.7), "host hess")
    cc = np.full(m.ncon, np.nan); assert_close(m.cons_nln(x, cc), ora.cons(x), "host cons")
    if m.ncon and m.nnzj:   # products and 32-bit structures through the host shims
        vh = np.random.default_rng(3).standard_normal(m.nvar); wh = np.random.default_rng(4).standard_normal(m.ncon)
        assert_close(m.jprod_nln(x, vh, np.full(m.ncon, np.nan)), ora.jprod(x, vh), "host jprod")
        assert_close(m.jtprod_nln(x, wh, np.full(m.nvar, np.nan)), ora.jtprod(x, wh), "host jtprod")
        assert_close(m.hprod(x, y, vh, np.full(m.nvar, np.nan), obj_weight=0.7), ora.hprod(x, y, vh, 0.7), "host hprod")
        r32, c32 = np.zeros(m.nnzj, dtype=np.int32), np.zeros(m.nnzj, dtype=np.int32)
        m.jac_structure(r32, c32)
        assert np.array_equal(r32.astype(np.int64), jr) and np.array_equal(c32.astype(np.int64), jc)
    r32, c32 = np.zeros(m.nnzh, dtype=np.int32), np.zeros(m.nnzh, dtype=np.int32)
    m.hess_structure(r32, c32)
    assert np.array_equal(r32.astype(np.int64), hr) and np.array_equal(c32.astype(np.int64), hc)
