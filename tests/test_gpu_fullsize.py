"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle cannot sweep
10^7 points in test time): windows of the big model against a small oracle model over the same
x-window, linearity of hess_coord! in (y, obj_weight), structure invariants, idempotence."""
import numpy as np
import pytest

from util import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lv_big(exa):
    import torch
    from examodels_jl_b200 import models as M
    N = 10_000_000
    core = M.luksan_vlcek(N)
    m = exa.ExaModel(core)
    rng = np.random.default_rng(0)
    x = M.lv_x0(N) + 0.01 * rng.uniform(-1, 1, N)
    y = np.random.default_rng(1).standard_normal(N - 2)
    return dict(N=N, m=m, x=x, y=y, dx=torch.from_numpy(x).cuda(), dy=torch.from_numpy(y).cuda(), torch=torch)


def test_lv_1e7_dims(lv_big):
    m, N = lv_big["m"], lv_big["N"]
    assert (m.nvar, m.ncon, m.nnzj, m.nnzh) == (N, N - 2, 3 * (N - 2), 9 * N - 15)


def test_lv_1e7_windows_match_oracle(lv_big):
    """Constraint point i reads x[i..i+2] only, objective point i reads x[i-1..i]: the slots of points
    [s, s+K) equal those of a K-point model built over the window x[s : s+K+2]."""
    from examodels_jl_b200 import models as M
    from oracle.oracle_api import Oracle
    m, N, torch = lv_big["m"], lv_big["N"], lv_big["torch"]
    h = m.hess_coord(lv_big["dx"], lv_big["dy"], m.new(m.nnzh), obj_weight=0.5)
    j = m.jac_coord(lv_big["dx"], m.new(m.nnzj))
    c = m.cons_nln(lv_big["dx"], m.new(m.ncon))
    K = 2000
    small = M.luksan_vlcek(K + 2)
    ora = Oracle.from_core(small)
    for s in (0, 1234567, 4999999, N - 2 - K):      # 0-based first constraint point of the window
        xw = lv_big["x"][s: s + K + 2]
        yw = lv_big["y"][s: s + K]
        ref_h = ora.hess_coord(xw, yw, 0.5)
        assert_close(h[6 * s: 6 * (s + K)].cpu().numpy(), ref_h[: 6 * K], f"hess con window {s}")
        # objective points: global point q (0-based, i = q+2) reads x[q], x[q+1]; window points q = s .. s+K
        o2 = 6 * (N - 2)
        assert_close(h[o2 + 3 * s: o2 + 3 * (s + K + 1)].cpu().numpy(), ref_h[6 * K: 6 * K + 3 * (K + 1)], f"hess obj window {s}")
        assert_close(j[3 * s: 3 * (s + K)].cpu().numpy(), ora.jac_coord(xw), f"jac window {s}")
        assert_close(c[s: s + K].cpu().numpy(), ora.cons(xw), f"cons window {s}")


def test_lv_1e7_hess_linearity_and_idempotence(lv_big):
    m, torch = lv_big["m"], lv_big["torch"]
    dx, dy = lv_big["dx"], lv_big["dy"]
    y2 = torch.from_numpy(np.random.default_rng(5).standard_normal(m.ncon)).cuda()
    h1 = m.hess_coord(dx, dy, m.new(m.nnzh), obj_weight=1.0)
    h1b = m.hess_coord(dx, dy, m.new(m.nnzh).fill_(float("nan")), obj_weight=1.0)
    assert torch.equal(h1, h1b)                                             # deterministic, fully overwritten
    h2 = m.hess_coord(dx, y2, m.new(m.nnzh), obj_weight=-0.25)
    h3 = m.hess_coord(dx, 2.0 * dy + 3.0 * y2, m.new(m.nnzh), obj_weight=2.0 - 0.75)
    lin = 2.0 * h1 + 3.0 * h2
    err = (h3 - lin).abs().max().item()
    scale = lin.abs().max().item()
    assert err <= 1e-10 * scale, (err, scale)
    ho = m.hess_coord(dx, None, m.new(m.nnzh), obj_weight=1.0)              # objective-only form
    assert ho[: 6 * (lv_big["N"] - 2)].abs().max().item() == 0.0


def test_lv_1e7_structure_invariants(lv_big):
    m, N, torch = lv_big["m"], lv_big["N"], lv_big["torch"]
    r, c = m.new(m.nnzh, torch.int32), m.new(m.nnzh, torch.int32)
    m.hess_structure(r, c)
    assert bool((r >= c).all()) and int(c.min()) == 1 and int(r.max()) == N
    # closed form from the known-answer layout: constraint point i (1-based) -> rows (i+1,i+2,i+2,i+2,i,i+1)
    i = torch.arange(1, N - 1, device="cuda", dtype=torch.int32)
    rows = torch.stack([i + 1, i + 2, i + 2, i + 2, i, i + 1], dim=1).reshape(-1)
    cols = torch.stack([i + 1, i + 2, i + 1, i + 1, i, i], dim=1).reshape(-1)
    assert torch.equal(r[: 6 * (N - 2)], rows) and torch.equal(c[: 6 * (N - 2)], cols)
    jr, jc = m.new(m.nnzj, torch.int64), m.new(m.nnzj, torch.int64)
    m.jac_structure(jr, jc)
    i64 = i.to(torch.int64)
    assert torch.equal(jr, i64.repeat_interleave(3))
    assert torch.equal(jc, torch.stack([i64 + 1, i64 + 2, i64], dim=1).reshape(-1))


def test_lv_1e7_grad_and_obj_consistency(lv_big):
    """grad! against central differences of obj along random directions (a checksum of the whole vector)."""
    m, torch = lv_big["m"], lv_big["torch"]
    dx = lv_big["dx"]
    g = m.grad(dx, m.new(m.nvar))
    d = torch.from_numpy(np.random.default_rng(9).standard_normal(m.nvar)).cuda()
    eps = 1e-6
    fd = (m.obj(dx + eps * d) - m.obj(dx - eps * d)) / (2 * eps)
    gd = float((g * d).sum())
    assert abs(fd - gd) <= 1e-6 * max(abs(gd), 1.0), (fd, gd)
    # windows of the dense gradient against the oracle (interior variables of a window see both neighbours)
    from examodels_jl_b200 import models as M
    from oracle.oracle_api import Oracle
    K = 1000
    ora = Oracle.from_core(M.luksan_vlcek(K))
    for s in (0, 7654321):
        ref = ora.grad(lv_big["x"][s: s + K])
        lo = 0 if s == 0 else 1
        assert_close(g[s + lo: s + K - 1].cpu().numpy(), ref[lo: K - 1], f"grad window {s}")


@pytest.mark.parametrize("which", ["lv_1e7", "rocket_1e6", "opf_10k", "family_32x1e6"])
def test_baseline_configs_full_size_direct(exa, which):
    """BASELINE.json configs 2-5 at their FULL sizes, every value callback against the oracle run with all host threads
    (a few seconds each), plus the integer structure where it fits comfortably in host memory."""
    import torch
    from examodels_jl_b200 import models as M
    from oracle.oracle_api import Oracle
    core = {"lv_1e7": lambda: M.luksan_vlcek(10_000_000), "rocket_1e6": lambda: M.goddard_rocket(1_000_000),
            "opf_10k": lambda: M.ac_power(M.synthetic_power_data()), "family_32x1e6": lambda: M.pattern_family(1_000_000, 32)}[which]()
    ora = Oracle.from_core(core)
    ora.set_threads(Oracle.max_threads())
    m = exa.ExaModel(core)
    meta = core.meta()
    x = meta["x0"] + 0.01 * np.random.default_rng(0).uniform(-1, 1, m.nvar)
    y = np.random.default_rng(1).standard_normal(m.ncon)
    dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    nan = float("nan")
    assert_close(m.hess_coord(dx, dy, m.new(m.nnzh).fill_(nan), obj_weight=0.5).cpu().numpy(), ora.hess_coord(x, y, 0.5), "hess")
    assert_close(m.jac_coord(dx, m.new(m.nnzj).fill_(nan)).cpu().numpy(), ora.jac_coord(x), "jac")
    assert_close(m.grad(dx, m.new(m.nvar).fill_(nan)).cpu().numpy(), ora.grad(x), "grad")
    assert_close(m.cons_nln(dx, m.new(m.ncon).fill_(nan)).cpu().numpy(), ora.cons(x), "cons")
    ref = ora.obj(x)
    assert abs(m.obj(dx) - ref) <= 1e-10 * max(1.0, abs(ref))
    if which in ("rocket_1e6", "opf_10k"):
        hr, hc = ora.hess_structure(); jr, jc = ora.jac_structure()
        r, c = m.new(m.nnzh, torch.int64), m.new(m.nnzh, torch.int64); m.hess_structure(r, c)
        assert np.array_equal(r.cpu().numpy(), hr) and np.array_equal(c.cpu().numpy(), hc)
        r, c = m.new(m.nnzj, torch.int32), m.new(m.nnzj, torch.int32); m.jac_structure(r, c)
        assert np.array_equal(r.cpu().numpy().astype(np.int64), jr) and np.array_equal(c.cpu().numpy().astype(np.int64), jc)


def test_lv_3e8_slot_numbers_beyond_int32(exa):
    """Maximum-size edge case: LV N = 3x10^8 has nnzh = 2.7x10^9 > 2^31 output slots (21.6 GB of values).  Windows of the
    result -- including the last one, whose slot numbers need 64-bit arithmetic -- against a small oracle model over the
    same x-window; structure of the last constraint points in closed form."""
    import torch
    from examodels_jl_b200 import models as M
    from oracle.oracle_api import Oracle
    N = 300_000_000
    free, _ = torch.cuda.mem_get_info()
    if free < 60 * 2 ** 30:
        pytest.skip("needs ~50 GB of free device memory")
    core = M.luksan_vlcek(N)
    m = exa.ExaModel(core)
    assert m.nnzh == 9 * N - 15 and m.nnzh > 2 ** 31
    g = torch.Generator(device="cuda").manual_seed(3)
    i = torch.arange(1, N + 1, device="cuda", dtype=torch.int32)
    dx = torch.where(i % 2 == 1, -1.2, 1.0).to(torch.float64) + 0.01 * (2.0 * torch.rand(N, device="cuda", dtype=torch.float64, generator=g) - 1.0)
    del i
    dy = torch.randn(N - 2, device="cuda", dtype=torch.float64, generator=g)
    h = m.hess_coord(dx, dy, m.new(m.nnzh), obj_weight=0.5)
    j = m.jac_coord(dx, m.new(m.nnzj))
    gr = m.grad(dx, m.new(m.nvar))
    K = 1500
    ora = Oracle.from_core(M.luksan_vlcek(K + 2))
    o2 = 6 * (N - 2)
    for s in (0, 123_456_789, 250_000_001, N - 2 - K):
        xw, yw = dx[s: s + K + 2].cpu().numpy(), dy[s: s + K].cpu().numpy()
        ref_h = ora.hess_coord(xw, yw, 0.5)
        assert_close(h[6 * s: 6 * (s + K)].cpu().numpy(), ref_h[: 6 * K], f"hess con window {s}")
        assert_close(h[o2 + 3 * s: o2 + 3 * (s + K + 1)].cpu().numpy(), ref_h[6 * K: 6 * K + 3 * (K + 1)], f"hess obj window {s}")
        assert_close(j[3 * s: 3 * (s + K)].cpu().numpy(), ora.jac_coord(xw), f"jac window {s}")
        ref_g = ora.grad(xw)
        lo = 0 if s == 0 else 1
        hi = K + 2 if s == N - 2 - K else K + 1
        assert_close(gr[s + lo: s + hi].cpu().numpy(), ref_g[lo: hi], f"grad window {s}")
    assert bool(torch.isfinite(h[-3 * K:]).all())
    del h, j, gr
    r, c = m.new(m.nnzh, torch.int32), m.new(m.nnzh, torch.int32)
    m.hess_structure(r, c)
    k = torch.arange(N - 2 - K + 1, N - 1, device="cuda", dtype=torch.int32)      # the last K constraint points (1-based i)
    rows = torch.stack([k + 1, k + 2, k + 2, k + 2, k, k + 1], dim=1).reshape(-1)
    cols = torch.stack([k + 1, k + 2, k + 1, k + 1, k, k], dim=1).reshape(-1)
    assert torch.equal(r[6 * (N - 2 - K): 6 * (N - 2)], rows) and torch.equal(c[6 * (N - 2 - K): 6 * (N - 2)], cols)
    assert int(r[-1]) == N and int(c[-1]) == N - 1     # last objective slot: (i, i-1) at i = N
