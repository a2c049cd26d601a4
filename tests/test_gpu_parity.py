"""GPU parity: every callback of the CUDA path (through the C ABI) against the CPU oracle on the
same seeded inputs.  Integer structure is compared with ==, FP64 values at 1e-10 (tests/util.py).

Model set follows the reference's own parity harness (test/NLPTest/NLPTest.jl:11-19,48-114):
LV (plain, both add orders), LV with augmentation + 2-D blocks (test/NLPTest/luksan.jl), the
AC-OPF pattern set (test/NLPTest/power.jl), plus the BASELINE configs' rocket and 32-pattern family.
"""
import numpy as np
import pytest

from util import assert_close, inputs

pytestmark = pytest.mark.gpu


def _models():
    from examodels_jl_b200 import models as M
    return {
        "lv3": lambda: M.luksan_vlcek(3),
        "lv20": lambda: M.luksan_vlcek(20),
        "lv100_bench": lambda: M.luksan_vlcek(100, order="bench"),
        "lv100_guide": lambda: M.luksan_vlcek(100, order="guide"),
        "lv1000": lambda: M.luksan_vlcek(1000),
        "lv_aug_20x1": lambda: M.luksan_vlcek_aug(20, 1),
        "lv_aug_20x3": lambda: M.luksan_vlcek_aug(20, 3),
        "opf_small": lambda: M.ac_power(M.synthetic_power_data(30, 41, 6, seed=5)),
        "opf_300": lambda: M.ac_power(M.synthetic_power_data(300, 420, 70, seed=2)),
        "rocket_50": lambda: M.goddard_rocket(50),
        "family_1000": lambda: M.pattern_family(1000, 32),
        "params": lambda: M.parametric(200),
        "all_ops_0": lambda: M.all_ops(64, 0),
        "all_ops_1": lambda: M.all_ops(64, 1),
        "all_ops_2": lambda: M.all_ops(64, 2),
        "all_ops_3_special": lambda: M.all_ops(64, 3),   # SpecialFunctions extension (ext/functionlist.jl)
    }


@pytest.fixture(scope="module")
def torch_():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


@pytest.mark.parametrize("name", list(_models().keys()))
def test_callbacks_match_oracle(exa, torch_, name):
    from oracle.oracle_api import Oracle
    torch = torch_
    core = _models()[name]()
    ora = Oracle.from_core(core)
    m = exa.ExaModel(core)
    assert (m.nvar, m.ncon, m.nnzj, m.nnzh) == (ora.nvar, ora.ncon, ora.nnzj, ora.nnzh)
    x, y = inputs(core)
    dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()

    # structure: bit-for-bit, int64 and int32
    jr, jc = ora.jac_structure()
    hr, hc = ora.hess_structure()
    for dt in (torch.int64, torch.int32):
        r, c = m.new(m.nnzj, dt), m.new(m.nnzj, dt)
        m.jac_structure(r, c)
        assert np.array_equal(r.cpu().numpy().astype(np.int64), jr), "jac rows"
        assert np.array_equal(c.cpu().numpy().astype(np.int64), jc), "jac cols"
        r, c = m.new(m.nnzh, dt), m.new(m.nnzh, dt)
        m.hess_structure(r, c)
        assert np.array_equal(r.cpu().numpy().astype(np.int64), hr), "hess rows"
        assert np.array_equal(c.cpu().numpy().astype(np.int64), hc), "hess cols"

    # values (outputs are pre-filled with garbage: the callee must fully define them)
    ref_obj = ora.obj(x)
    got_obj = m.obj(dx)
    assert abs(got_obj - ref_obj) <= 1e-10 * max(abs(ref_obj), 1.0), (got_obj, ref_obj)
    g = m.new(m.nvar).fill_(float("nan"))
    assert_close(m.grad(dx, g).cpu().numpy(), ora.grad(x), "grad")
    c = m.new(m.ncon).fill_(float("nan"))
    assert_close(m.cons_nln(dx, c).cpu().numpy(), ora.cons(x), "cons")
    j = m.new(m.nnzj).fill_(float("nan"))
    assert_close(m.jac_coord(dx, j).cpu().numpy(), ora.jac_coord(x), "jac")
    h = m.new(m.nnzh).fill_(float("nan"))
    assert_close(m.hess_coord(dx, dy, h, obj_weight=1.0).cpu().numpy(), ora.hess_coord(x, y, 1.0), "hess")
    h.fill_(float("nan"))
    assert_close(m.hess_coord(dx, dy, h, obj_weight=0.5).cpu().numpy(), ora.hess_coord(x, y, 0.5), "hess s=0.5")
    h.fill_(float("nan"))
    assert_close(m.hess_coord(dx, None, h, obj_weight=2.0).cpu().numpy(), ora.hess_coord(x, None, 2.0), "hess obj-only")

    # host-buffer shims (WrapperNLPModel role)
    hh = np.full(m.nnzh, np.nan)
    assert_close(m.hess_coord(x, y, hh, obj_weight=1.0), ora.hess_coord(x, y, 1.0), "host hess")
    gg = np.full(m.nvar, np.nan)
    assert_close(m.grad(x, gg), ora.grad(x), "host grad")
    assert abs(m.obj(x) - ref_obj) <= 1e-10 * max(abs(ref_obj), 1.0)
    st = m.stats()
    assert st["launches"] > 0


def test_cuda_graph_full_eval_matches_eager(exa, torch_):
    """obj + grad! + cons! + jac_coord! + hess_coord! captured in one CUDA graph give the eager results,
    also after x / y are updated in place."""
    from examodels_jl_b200 import models as M
    torch = torch_
    core = M.ac_power(M.synthetic_power_data(300, 420, 70, seed=2))
    m = exa.ExaModel(core)
    x, y = inputs(core)
    dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    od, g, c, j, h = m.new(1), m.new(m.nvar), m.new(m.ncon), m.new(m.nnzj), m.new(m.nnzh)
    gr = m.capture_full_eval(dx, dy, od, g, c, j, h, obj_weight=0.5)
    for seed in (3, 4):
        x2, y2 = inputs(core, seed)
        dx.copy_(torch.from_numpy(x2)); dy.copy_(torch.from_numpy(y2))
        for t in (od, g, c, j, h):
            t.fill_(float("nan"))
        gr.replay()
        torch.cuda.synchronize()
        assert od.item() == m.obj(dx)
        assert torch.equal(g, m.grad(dx, m.new(m.nvar)))
        assert torch.equal(c, m.cons_nln(dx, m.new(m.ncon)))
        assert torch.equal(j, m.jac_coord(dx, m.new(m.nnzj)))
        assert torch.equal(h, m.hess_coord(dx, dy, m.new(m.nnzh), obj_weight=0.5))


def test_cuda_graph_of_the_fused_evaluation(exa, torch_):
    """`exb_eval` (one sweep + finishing steps) captured as one CUDA graph replays to the oracle's values after x / y change."""
    from examodels_jl_b200 import models as M
    from oracle.oracle_api import Oracle
    torch = torch_
    for core in (M.ac_power(M.synthetic_power_data(300, 420, 70, seed=2)), M.luksan_vlcek(5000)):
        m, ora = exa.ExaModel(core), Oracle.from_core(core)
        x, y = inputs(core)
        dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
        od, g, c, j, h = m.new(1), m.new(m.nvar), m.new(m.ncon), m.new(m.nnzj), m.new(m.nnzh)
        gr = m.capture_fused_eval(dx, dy, od, g, c, j, h, obj_weight=0.5)
        x2, y2 = inputs(core, 5)
        dx.copy_(torch.from_numpy(x2)); dy.copy_(torch.from_numpy(y2))
        for t in (od, g, c, j, h):
            t.fill_(float("nan"))
        gr.replay()
        torch.cuda.synchronize()
        assert abs(od.item() - ora.obj(x2)) <= 1e-10 * max(1.0, abs(ora.obj(x2)))
        assert_close(g.cpu().numpy(), ora.grad(x2), "graph grad")
        assert_close(c.cpu().numpy(), ora.cons(x2), "graph cons")
        assert_close(j.cpu().numpy(), ora.jac_coord(x2), "graph jac")
        assert_close(h.cpu().numpy(), ora.hess_coord(x2, y2, 0.5), "graph hess")


def test_plain_c_host_through_the_abi(exa, tmp_path):
    """The boundary is usable without Python: a C program (tests/c_host_example.c) linked against libexa_b200.so
    builds LV N=5000 from an IR file and evaluates hess_coord! / obj with host buffers."""
    import os
    import subprocess
    from examodels_jl_b200 import models as M
    from oracle.oracle_api import Oracle
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(root, "examodels.jl_b200", "csrc")
    exe = str(tmp_path / "c_host_example")
    subprocess.check_call(["gcc", "-O1", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "c_host_example.c"),
                           "-o", exe, "-L", csrc, "-lexa_b200", "-Wl,-rpath," + csrc])
    core = M.luksan_vlcek(5000)
    ir, bufs = core.to_ir()
    assert not bufs
    x, y = inputs(core)
    (tmp_path / "m.ir").write_bytes(ir); x.tofile(tmp_path / "x.bin"); y.tofile(tmp_path / "y.bin")
    out = subprocess.check_output([exe, str(tmp_path / "m.ir"), str(tmp_path / "x.bin"), str(tmp_path / "y.bin"), "0.5"], text=True)
    nnzh, s, w, obj = out.split()
    ora = Oracle.from_core(core)
    h = ora.hess_coord(x, y, 0.5)
    wts = (np.arange(h.size) % 97) + 1.0
    assert int(nnzh) == ora.nnzh
    scale = np.abs(h).sum()
    assert abs(float(s) - h.sum()) <= 1e-10 * scale and abs(float(w) - (h * wts).sum()) <= 1e-10 * (np.abs(h) * wts).sum()
    assert abs(float(obj) - ora.obj(x)) <= 1e-10 * abs(ora.obj(x))


def test_per_callback_timing(exa, torch_):
    """The TimedNLPModel role: CUDA-event timing per callback, call counts, reset."""
    from examodels_jl_b200 import models as M
    torch = torch_
    core = M.luksan_vlcek(200_000)
    m = exa.ExaModel(core)
    x, y = inputs(core)
    dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    h, j, g, c = m.new(m.nnzh), m.new(m.nnzj), m.new(m.nvar), m.new(m.ncon)
    m.hess_coord(dx, dy, h); m.jac_coord(dx, j); m.grad(dx, g); m.cons_nln(dx, c); m.obj(dx)   # tuning calls, untimed
    m.set_timing(True)
    for _ in range(3):
        m.hess_coord(dx, dy, h); m.jac_coord(dx, j)
    m.grad(dx, g); m.cons_nln(dx, c); m.obj(dx)
    t = m.timings(reset=True)
    assert [t[k]["calls"] for k in ("obj", "grad", "cons", "jac", "hess")] == [1, 1, 1, 3, 3]
    assert all(t[k]["ms"] > 0 for k in ("obj", "grad", "cons", "jac", "hess")) and t["hprod"]["calls"] == 0
    assert t["hess"]["ms"] < 50.0
    m.set_timing(False)
    m.hess_coord(dx, dy, h)
    assert m.timings()["hess"]["calls"] == 0


@pytest.mark.parametrize("which", ["lv", "lv_aug", "rocket"])
def test_pipelined_host_shims(exa, torch_, which):
    """exb_host_hess / exb_host_jac with page-locked caller buffers: windowed launches on one stream, D2H of each
    window's per-pattern slices on a second one (csrc/exb_runtime.cpp host_coo_pipelined).  Same values as the oracle;
    pageable buffers keep the single-launch path.  `lv_aug` has augmentation rows, so y is uploaded whole."""
    from examodels_jl_b200 import models as M
    from oracle.oracle_api import Oracle
    torch = torch_
    core = {"lv": lambda: M.luksan_vlcek(400_000), "lv_aug": lambda: M.luksan_vlcek_aug(80_000, 3),
            "rocket": lambda: M.goddard_rocket(60_000)}[which]()
    m, ora = exa.ExaModel(core), Oracle.from_core(core)
    x, y = inputs(core)
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()   # noqa: E731
    xp, yp = pin(x), pin(y)
    hp, jp = pin(np.full(m.nnzh, np.nan)), pin(np.full(m.nnzj, np.nan))
    m.hess_coord(xp, yp, hp, obj_weight=0.5); m.jac_coord(xp, jp)          # first calls tune the kernels (plain path)
    ref_h, ref_j = ora.hess_coord(x, y, 0.5), ora.jac_coord(x)
    assert_close(hp, ref_h, "host hess, tuning call"); assert_close(jp, ref_j, "host jac, tuning call")
    for rep in range(2):
        hp[:] = np.nan; jp[:] = np.nan
        l0 = m.stats()["launches"]
        m.hess_coord(xp, yp, hp, obj_weight=0.5)
        l1 = m.stats()["launches"]
        m.jac_coord(xp, jp)
        l2 = m.stats()["launches"]
        assert l1 - l0 > 1 and l2 - l1 > 1, "pinned buffers should take the windowed path"
        assert_close(hp, ref_h, "pipelined host hess"); assert_close(jp, ref_j, "pipelined host jac")
    hp[:] = np.nan
    m.hess_coord(xp, None, hp, obj_weight=2.0)
    assert_close(hp, ora.hess_coord(x, None, 2.0), "pipelined host hess, objective only")
    hh = np.full(m.nnzh, np.nan)
    l0 = m.stats()["launches"]
    assert_close(m.hess_coord(x, y, hh, obj_weight=0.5), ref_h, "pageable host hess")
    assert m.stats()["launches"] - l0 == 1


@pytest.mark.parametrize("which", ["lv_bench", "lv_guide_ragged", "only_objective", "lv_sharded"])
def test_persistent_hessian_kernel(exa, torch_, which, monkeypatch):
    """The experimental persistent form of the Hessian kernel (x / y windows of the next tile prefetched into shared memory
    with cp.async, csrc/exb_device.cuh exb_hessp_body; opt-in, measured slower than the classic kernel) forced on: same values as the oracle, for every obj_weight / y form,
    ragged last tiles and a sharded handle; and the classic form forced on gives bitwise the same vector."""
    from examodels_jl_b200 import models as M
    from oracle.oracle_api import Oracle
    from edge_models import EDGE
    torch = torch_
    core = {"lv_bench": lambda: M.luksan_vlcek(200_003), "lv_guide_ragged": lambda: M.luksan_vlcek(1029, order="guide"),
            "only_objective": EDGE["only_objective"], "lv_sharded": lambda: M.luksan_vlcek(70_001)}[which]()
    ora = Oracle.from_core(core)
    x, y = inputs(core)
    dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    kw = dict(rank=1, world=3) if which == "lv_sharded" else {}
    monkeypatch.setenv("EXB_TUNE_PERSISTENT", "1")          # opt-in: generate the persistent kernel next to the classic one
    monkeypatch.setenv("EXB_TUNE_FORCE_PERSISTENT", "1")
    mp_ = exa.ExaModel(core, **kw)
    assert mp_.kernel_choice("hess")["persistent"]
    monkeypatch.setenv("EXB_TUNE_FORCE_PERSISTENT", "0")
    mc = exa.ExaModel(core, **kw)
    assert not mc.kernel_choice("hess")["persistent"]
    if kw:
        ora.set_shard(1, 3)
    for yy, w in ((dy, 1.0), (dy, 0.5), (None, 2.0)):
        hp = mp_.hess_coord(dx, yy, mp_.new(mp_.nnzh).fill_(float("nan")), obj_weight=w)
        hc = mc.hess_coord(dx, yy, mc.new(mc.nnzh).fill_(float("nan")), obj_weight=w)
        assert torch.equal(torch.nan_to_num(hp, nan=-7.0), torch.nan_to_num(hc, nan=-7.0)), "persistent and classic kernels differ"
        ref = ora.hess_coord(x, None if yy is None else y, w)
        got = hp.cpu().numpy()
        if kw:   # a sharded handle writes only its slices
            mine = ~np.isnan(got)
            assert mine.any() and not mine.all()
            assert_close(got[mine], ref[mine], f"persistent hess shard (w={w})")
        else:
            assert_close(got, ref, f"persistent hess (w={w})")


def test_sharded_host_shims_upload_only_what_the_shard_reads(exa, torch_):
    """A sharded handle's host entry points copy the part of x its points can read (shifts of a range iterator are known
    to the plan), its own rows of y and the slices it wrote -- and still match the oracle."""
    from examodels_jl_b200 import models as M
    from oracle.oracle_api import Oracle
    torch = torch_
    core = M.luksan_vlcek(400_000)
    ora = Oracle.from_core(core)
    x, y = inputs(core)
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()   # noqa: E731
    xp, yp = pin(x), pin(y)
    ref = ora.hess_coord(x, y, 0.5)
    gref = ora.grad(x)
    got = np.full(ora.nnzh, np.nan)
    tot_h2d = tot_d2h = 0
    for r in range(4):
        m = exa.ExaModel(core, rank=r, world=4)
        hp = pin(np.full(m.nnzh, np.nan))
        m.hess_coord(xp, yp, hp, obj_weight=0.5)       # tuning call (plain path)
        hp[:] = np.nan
        m.hess_coord(xp, yp, hp, obj_weight=0.5)       # windowed path
        h2d, d2h = m.host_bytes()
        assert h2d <= 8 * (m.nvar // 4 + 8 + m.ncon // 4 + 8) and d2h <= 8 * (m.nnzh // 4 + 16), (h2d, d2h)
        tot_h2d += h2d; tot_d2h += d2h
        mine = ~np.isnan(hp)
        assert not (mine & ~np.isnan(got)).any()
        got[mine] = hp[mine]
        g = np.full(m.nvar, np.nan)
        m.grad(x, g)                                    # pageable buffers, partial x upload as well
        # the gradient of a shift-indexed objective is owner-computed per VARIABLE: exact on the owned range, zero elsewhere
        lo, hi = m.owned()
        assert (lo, hi) == (m.nvar * r // 4, m.nvar * (r + 1) // 4)
        assert_close(g[lo:hi], gref[lo:hi], f"host grad shard {r} (owned variables)")
        assert not g[:lo].any() and not g[hi:].any()
    assert d2h > 0 and tot_d2h == 8 * ora.nnzh
    assert_close(got, ref, "sharded host hess, assembled")


def test_reference_ipopt_solution_is_a_kkt_point_on_the_gpu(exa, torch_):
    """The reference-produced fixture of docs/src/develop.md:84-105 (Ipopt solution + multipliers of LV N=10) through the CUDA
    path: constraints vanish, grad f + J' lambda = 0, the reduced Lagrangian Hessian is positive definite."""
    import json
    import os
    from examodels_jl_b200 import models as M
    torch = torch_
    sol = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "known_answers.json")))["lv10_ipopt_solution"]
    x, lam = np.array(sol["x"]), np.array(sol["multipliers"])
    m = exa.ExaModel(M.luksan_vlcek(10, order="guide"))
    dx, dl = torch.from_numpy(x).cuda(), torch.from_numpy(lam).cuda()
    jr, jc = m.new(m.nnzj, torch.int64), m.new(m.nnzj, torch.int64); m.jac_structure(jr, jc)
    J = np.zeros((m.ncon, m.nvar)); np.add.at(J, (jr.cpu().numpy() - 1, jc.cpu().numpy() - 1), m.jac_coord(dx, m.new(m.nnzj)).cpu().numpy())
    g = m.grad(dx, m.new(m.nvar)).cpu().numpy()
    c = m.cons_nln(dx, m.new(m.ncon)).cpu().numpy()
    assert np.abs(c).max() < 1e-10 and np.abs(g + J.T @ lam).max() < 2e-8
    jt = m.jtprod_nln(dx, dl, m.new(m.nvar)).cpu().numpy()          # the fused J' v kernel gives the same residual
    assert np.abs(g + jt).max() < 2e-8
    hr, hc = m.new(m.nnzh, torch.int64), m.new(m.nnzh, torch.int64); m.hess_structure(hr, hc)
    L = np.zeros((m.nvar, m.nvar)); np.add.at(L, (hr.cpu().numpy() - 1, hc.cpu().numpy() - 1), m.hess_coord(dx, dl, m.new(m.nnzh)).cpu().numpy())
    H = L + np.tril(L, -1).T
    Z = np.linalg.svd(J)[2][m.ncon:].T
    assert np.linalg.eigvalsh(Z.T @ H @ Z).min() > 0.0


def test_cuda_path_replays_the_reference_ipopt_logs(exa, torch_):
    """tests/test_oracle_pins.py::test_oracle_replays_the_reference_ipopt_logs_digit_for_digit, through the CUDA path: the
    three Ipopt runs printed by the reference's documentation build (docs/src/parameters.md) are reproduced digit for digit
    from exb_obj / exb_grad / exb_cons / exb_jac / exb_hess and the two structure kernels, with the parameters updated in
    place (exb_set_params) between the runs."""
    import json
    import os
    from examodels_jl_b200 import models as M
    from util import check_ipopt_run, newton_kkt_replay
    torch = torch_
    logs = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ipopt_logs.json")))
    core = M.luksan_vlcek_param(10)
    m = exa.ExaModel(core)
    jr, jc = m.new(m.nnzj, torch.int64), m.new(m.nnzj, torch.int64); m.jac_structure(jr, jc)
    hr, hc = m.new(m.nnzh, torch.int64), m.new(m.nnzh, torch.int64); m.hess_structure(hr, hc)
    jr, jc, hr, hc = (t.cpu().numpy() - 1 for t in (jr, jc, hr, hc))
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()   # noqa: E731

    def jac(x):
        J = np.zeros((m.ncon, m.nvar)); np.add.at(J, (jr, jc), m.jac_coord(dev(x), m.new(m.nnzj)).cpu().numpy()); return J

    def hess(x, lam, sigma):
        L = np.zeros((m.nvar, m.nvar)); np.add.at(L, (hr, hc), m.hess_coord(dev(x), dev(lam), m.new(m.nnzh), obj_weight=sigma).cpu().numpy())
        return L + np.tril(L, -1).T
    cb = dict(obj=lambda x: m.obj(dev(x)), grad=lambda x: m.grad(dev(x), m.new(m.nvar)).cpu().numpy(),
              cons=lambda x: m.cons_nln(dev(x), m.new(m.ncon)).cpu().numpy(), jac=jac, hess=hess)
    for run in logs["runs"]:
        assert (m.nnzj, m.nnzh) == (run["nnzj"], run["nnzh"])
        m.set_params(run["theta"])
        rows, final, _, _ = newton_kkt_replay(cb, core.meta()["x0"], len(run["iterations"]) - 1)
        check_ipopt_run(run, rows, final)


def test_split_hessian_pattern(exa, torch_, monkeypatch):
    """Opt-in (EXB_TUNE_SPLIT_NS): a pattern with many second-order slots per point is evaluated by two launch entries that keep
    one half of its slots each (csrc/exb_device.cuh ExbSplit / exb_store_rows; measured slower on the rocket, hence opt-in):
    same vector as the whole-pattern kernel, bit for bit, incl. ragged tiles and a sharded handle."""
    from examodels_jl_b200 import models as M
    torch = torch_
    core = M.goddard_rocket(50)
    x, y = inputs(core)
    dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    whole = exa.ExaModel(core)
    ref = whole.hess_coord(dx, dy, whole.new(whole.nnzh).fill_(float("nan")), obj_weight=0.7)
    monkeypatch.setenv("EXB_TUNE_SPLIT_NS", "32")
    assert "ExbSplit<P2, 0, 24>, ExbSplit<P2, 24, 47>" in exa.Plan(core).source()
    for kw in ({}, dict(rank=1, world=2)):
        split = exa.ExaModel(core, **kw)
        got = split.hess_coord(dx, dy, split.new(split.nnzh).fill_(float("nan")), obj_weight=0.7)
        mine = ~torch.isnan(got)
        assert mine.any() and (kw or mine.all()) and torch.equal(got[mine], ref[mine])
