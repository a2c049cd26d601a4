"""CPU-side checks of the product's host logic: the plan (probe, counters, generated source)
against the oracle, the C-ABI surface, IR validation, and failing loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import examodels_jl_b200 as E
from examodels_jl_b200 import models as M
from examodels_jl_b200 import backend as B
from oracle.oracle_api import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

MODELS = {
    "lv_bench": lambda: M.luksan_vlcek(50, order="bench"),
    "lv_guide": lambda: M.luksan_vlcek(50, order="guide"),
    "lv_aug": lambda: M.luksan_vlcek_aug(9, 3),
    "opf": lambda: M.ac_power(M.synthetic_power_data(30, 41, 6, seed=5)),
    "rocket": lambda: M.goddard_rocket(10),
    "family": lambda: M.pattern_family(10, 32),
}


@pytest.fixture(scope="module", autouse=True)
def _built():
    E.build_library()


@pytest.mark.parametrize("name", list(MODELS))
def test_plan_matches_oracle(name):
    core = MODELS[name]()
    p, o = E.Plan(core), Oracle.from_core(core)
    for a in ("nvar", "ncon", "nnzj", "nnzh", "nobj", "nnzg", "nconaug", "npar"):
        assert getattr(p, a) == getattr(o, a), a
    assert p.npatterns() == o.npatterns()
    for k in range(p.npatterns()):
        assert p.pattern_info(k) == o.pattern_info(k)
        assert np.array_equal(p.comp(k, 1), o.comp(k, 1))
        assert np.array_equal(p.comp(k, 2), o.comp(k, 2))


def test_generated_source_is_model_size_independent():
    a, b = E.Plan(M.luksan_vlcek(100)), E.Plan(M.luksan_vlcek(10_000))
    assert a.source() == b.source() and a.module_path() == b.module_path()
    src = a.source()
    for kern in ("exb_hess_g0", "exb_jac_g0", "exb_ggrad_g0", "exb_cons_g0", "exb_obj_g0", "exb_hstruct64_g0", "exb_eval_g0", "exb_hessc_g0"):
        assert f'extern "C" __global__ void __launch_bounds__(EXB_BLOCK, EXB_MINB) {kern}' in src
    assert "sincos" in src and "struct P0" in src and "struct P1" in src


def _has(src, kernel):
    return f"EXB_MINB) {kernel}(const ExbGroup" in src


def test_gradient_kernel_choice(monkeypatch):
    """Three gradient forms, chosen per objective pattern at build time.  Slots that address x[t + const] (t a range value, or the
    point number when an AoS iterator carries an iota column): (a) a light body that reads no iterator data is re-evaluated
    once per slot by the per-variable owner-computes kernel (LV); (b) anything else goes to the tile kernel -- a block per tile
    of variables evaluates the points around it ONCE and gathers their slots in the reference's order.  (c) Objectives
    indexed through iterator data keep the slot + segmented-sum path of the reference (ext:310-336,691-697)."""
    lv = E.Plan(M.luksan_vlcek(50)).source()
    assert _has(lv, "exb_ggrad_g0") and not _has(lv, "exb_sgrad_g0") and not _has(lv, "exb_gradt_g0")
    # summation order for variable v: point v (slot of x[i]) before point v + 1 (slot of x[i-1]) = ascending slot number
    g1 = lv[lv.index("double g1("):]
    assert g1.index("s[1] : 0.0") < g1.index("s[0] : 0.0")
    fam = E.Plan(M.pattern_family(100, 8)).source()       # `i` column = 1..n: recognised from the data, never loaded
    assert _has(fam, "exb_gradt_g0") and not _has(fam, "exb_sgrad_g0") and not _has(fam, "exb_ggrad_g0")
    gg = fam[fam.rindex("void ggather("):]
    gg = gg[:gg.index("}")]
    assert gg.count("acc[0] +=") >= 1
    opf = E.Plan(M.ac_power(M.synthetic_power_data(30, 41, 6, seed=5))).source()
    assert _has(opf, "exb_gradt_g0")                        # generator cost over pg[g.i], g.i = 1..ngen: reads cost data -> tile form
    monkeypatch.setenv("EXB_NO_TGRAD", "1")
    fam1 = E.Plan(M.pattern_family(100, 8)).source()
    assert _has(fam1, "exb_sgrad_g0") and not _has(fam1, "exb_gradt_g0")
    monkeypatch.delenv("EXB_NO_TGRAD")
    monkeypatch.setenv("EXB_NO_IOTA", "1")
    fam2 = E.Plan(M.pattern_family(100, 8)).source()
    assert _has(fam2, "exb_sgrad_g0") and not _has(fam2, "exb_gradt_g0") and not _has(fam2, "exb_ggrad_g0")


def test_iota_columns_are_recognised_from_the_data(monkeypatch):
    """An integer field holding v0, v0 + 1, ... turns a data-indexed pattern into a shift-indexed one: index relations are decided
    at build time, the duplicate-free Hessian is fused, the column is never loaded.  A permuted column is left alone."""
    core = M.pattern_family(200, 8)
    p = E.Plan(core)
    src = p.source()
    assert p.tile_info()["fused"] and p.tile_info()["nnzh_unique"] == 3 * 202 - 3
    p0 = src[src.index("struct P0 {"):src.index("struct P1 {")]
    assert "exb_ld_i(" not in p0 and "exb_twice_if_eq(" not in p0 and "kg + 1LL" in p0
    # same structure and counts as without the hint
    monkeypatch.setenv("EXB_NO_IOTA", "1")
    q = E.Plan(core)
    assert not q.tile_info()["fused"] and "exb_ld_i(" in q.source()
    for k in range(p.npatterns()):
        assert p.pattern_info(k) == q.pattern_info(k)
    monkeypatch.delenv("EXB_NO_IOTA")
    import numpy as np
    core2 = M.pattern_family(200, 2)
    core2.patterns[0].itr.array["i"][:] = np.random.default_rng(0).permutation(200) + 1
    assert "exb_ld_i(" in E.Plan(core2).source()[:E.Plan(core2).source().index("struct P1 {")]


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "exa_b200.h")).read()
    decl = set(re.findall(r"\b(exb_[a-z0-9_]+)\s*\(", hdr))
    assert len(decl) >= 30
    lib = B.lib()
    for s in decl:
        assert hasattr(lib, s), f"{s} declared in include/exa_b200.h but not exported"
    assert lib.exb_abi_version() == 2


def test_malformed_ir_is_rejected():
    lib = B.lib()
    h = C.c_void_p()
    bad = b"\0" * 64
    assert lib.exb_plan_create(bad, C.c_size_t(len(bad)), None, C.byref(h)) == 3
    assert b"magic" in lib.exb_last_error()
    ir, _ = M.luksan_vlcek(10).to_ir()
    assert lib.exb_plan_create(ir[:-24], C.c_size_t(len(ir) - 24), None, C.byref(h)) == 3


def test_no_silent_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(E.ExbError):
        E.ExaModel(M.luksan_vlcek(10))
    # the raw ABI refuses as well
    lib = B.lib()
    ir, bufs = M.luksan_vlcek(10).to_ir()
    h = C.c_void_p()
    rc = lib.exb_create(ir, C.c_size_t(len(ir)), None, 0, None, C.byref(h))
    assert rc == 5 and not h.value


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "examodels.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".hpp", ".h")) and f != "exb_embed.cpp":
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle_api" not in txt and "exa_oracle" not in txt and "libexa_oracle" not in txt, f


def test_front_end_tree_shapes():
    from examodels_jl_b200 import graph as G
    c = E.ExaCore()
    x = c.add_var(5)
    d = G.DataSource()
    assert isinstance(x[d] ** 2, G.Node1) and (x[d] ** 2).op == "abs2"          # specialization.jl:198
    p3 = x[d] ** 3
    assert isinstance(p3, G.Node2) and p3.op == "^" and isinstance(p3.inner2, G.Val)   # :199
    assert (x[d] ** 1) is not None and isinstance(x[d] ** 1, G.Var)
    e = 3 * x[d] ** 3                                                              # 3*(x^3)
    assert isinstance(e, G.Node2) and e.op == "*" and e.inner1 == 3
    s = x[d] + x[d + 1] + x[d + 2]                                                 # left fold
    assert s.inner1.op == "+" and isinstance(s.inner2, G.Var)
    assert x[d] == x[d] and x[d + 1] != x[d]                                       # === on index expressions
    assert x[3].i == 3 and isinstance(x[d].i, G.Node2)                             # nlp.jl:908,922


def _generated(core):
    return E.Plan(core).source().split("generated by exb_plan")[1]


def test_index_equality_is_decided_at_build_time_for_affine_maps():
    """`i == j ? 2adj : adj` (src/hessian.jl:261-266) needs a run-time compare only when the two index expressions can
    coincide: affine maps of a range iterator are decided by the plan, data-driven indices are not."""
    import numpy as np
    from examodels_jl_b200.graph import sin

    def model(body, itr):
        c = E.ExaCore(); x = c.add_var(64, start=np.linspace(0.1, 0.9, 64))
        c.add_con(lambda i: body(x, i), itr)
        return c
    assert "exb_twice_if_eq" not in _generated(model(lambda x, i: x[i] * x[i + 1], range(1, 30)))          # never equal
    assert "exb_twice_if_eq" not in _generated(model(lambda x, i: sin(x[i] * x[2 * i + 5]), range(1, 25)))   # i = 2i + 5 has no solution in range
    assert "exb_twice_if_eq" in _generated(model(lambda x, i: x[i] * x[2 * i - 3], range(1, 30)))          # equal at i = 3
    assert "exb_twice_if_eq" not in _generated(model(lambda x, i: x[i] * x[2 * i - 3], range(4, 30)))      # ... which is outside this range
    d = np.zeros(5, dtype=np.dtype([("f", "i8"), ("t", "i8")])); d["f"] = [1, 2, 3, 4, 5]; d["t"] = [2, 2, 4, 4, 6]
    assert "exb_twice_if_eq" in _generated(model(lambda x, b: x[b.f] * x[b.t], d))                         # data: run-time compare


def test_persistent_kernel_is_opt_in_and_needs_windows(monkeypatch):
    lv = M.luksan_vlcek(50)
    assert "exb_hessp_g0" not in _generated(lv)
    monkeypatch.setenv("EXB_TUNE_PERSISTENT", "1")
    src = _generated(lv)
    assert "exb_hessp_g0" in src and "XLO = 0, XHI = 2" in src and "XLO = -1, XHI = 0" in src      # LV constraint / objective windows
    assert "exb_hessp_g0" not in _generated(M.luksan_vlcek_aug(10, 2))                             # data-indexed (product iterator): no window
    assert "exb_hessp_g0" in _generated(M.pattern_family(50, 8))                                   # iota column: as good as a range
    monkeypatch.setenv("EXB_NO_IOTA", "1")
    assert "exb_hessp_g0" not in _generated(M.pattern_family(50, 8))


def test_fixed_index_variables_bypass_the_window():
    """A variable at a fixed index (the rocket's step length) is read with ExbX*::ldc, everything else with ::ld."""
    src = _generated(M.goddard_rocket(20))
    assert "x.ldc(" in src and "x.ld(" in src
    assert "x.ldc(" not in _generated(M.luksan_vlcek(20))


def test_kernel_modules_are_cached_compressed(tmp_path, monkeypatch):
    """nvcc output is kept as `<hash>.cubin.gz` (the debug PTX text of -lineinfo is ~80 % of a cubin and compresses 6x);
    the decompressed image is a loadable sm_100a ELF with every callback kernel in it."""
    import gzip
    import subprocess
    monkeypatch.setenv("EXB_CACHE_DIR", str(tmp_path))
    p = E.Plan(M.luksan_vlcek(30))
    path = p.compile()
    assert path.startswith(str(tmp_path)) and not os.path.exists(path) and os.path.exists(path + ".gz")
    raw = gzip.open(path + ".gz", "rb").read()
    assert raw[:4] == b"\x7fELF" and len(raw) > 5 * os.path.getsize(path + ".gz") / 2
    (tmp_path / "m.cubin").write_bytes(raw)
    out = subprocess.run(["cuobjdump", "-res-usage", str(tmp_path / "m.cubin")], capture_output=True, text=True).stdout
    for k in ("exb_hess_g0", "exb_jac_g0", "exb_ggrad_g0", "exb_cons_g0", "exb_obj_g0", "exb_hprod_g0", "exb_eval_g0", "exb_hessc_g0"):
        assert f"Function {k}:" in out
    again = E.Plan(M.luksan_vlcek(31))          # same source: cache hit on the compressed module, nothing recompiled
    t0 = os.path.getmtime(path + ".gz")
    assert again.compile() == path and os.path.getmtime(path + ".gz") == t0
    monkeypatch.setenv("EXB_KEEP_CUBIN", "1")   # development: keep the raw module next to it
    q = E.Plan(M.luksan_vlcek(30, order="guide")).compile()
    assert os.path.exists(q) and os.path.exists(q + ".gz")


def test_kernel_pattern_lists_keep_objectives_out_of_the_constraint_kernels():
    """An objective pattern with a fixed-index variable (no tile / per-variable gradient form) goes to the slot kernel, never
    to exb_jac_g0 / exb_cons_g0 (regression: a dangling else once put it there and its gradient slot landed in jac[0])."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from edge_models import EDGE
    src = E.Plan(EDGE["single_points_and_constants"]()).source()

    def lst(kernel):
        line = next(ln for ln in src.splitlines() if f"EXB_MINB) {kernel}(const ExbGroup" in ln)
        return re.findall(r"P(\d+)", line[line.index("<"):line.index(">")])
    assert lst("exb_sgrad_g0") == ["0"] and lst("exb_obj_g0") == ["0"]
    assert "0" not in lst("exb_jac_g0") and "0" not in lst("exb_cons_g0")
    assert lst("exb_eval_g0") == [str(k) for k in range(6)]


def test_ir_validation_rejects_what_the_round_1_review_listed():
    """Unknown iterator kind / field type, DATA_SELF on an AoS iterator, DATA_FIELD on a range, a field that sticks out of its
    element, negative nvar: EXB_ERR_IR (3) with a message, never a crash or a silent misread."""
    import struct
    lib = B.lib()

    def create(words):
        ir = struct.pack(f"<{len(words)}q", *words)
        h = C.c_void_p()
        rc = lib.exb_plan_create(ir, C.c_size_t(len(ir)), None, C.byref(h))
        if rc == 0:
            lib.exb_plan_destroy(h)
        return rc, lib.exb_last_error().decode()

    def words_of(core):
        ir, _ = core.to_ir()
        return list(struct.unpack(f"<{len(ir) // 8}q", ir))
    lv = words_of(M.luksan_vlcek(10))
    assert create(lv)[0] == 0
    bad = list(lv); bad[2] = -5                      # nvar
    assert create(bad) == (3, "negative nvar / npar")
    bad = list(lv); bad[8] = 7                       # itr_kind of pattern 0 (header 6 words; kind, nitr, itr_kind)
    assert create(bad)[0] == 3 and "iterator kind" in create(bad)[1]
    d = np.zeros(4, dtype=np.dtype([("i", "i8"), ("a", "f8")])); d["i"] = [1, 2, 3, 4]
    c = E.ExaCore(); x = c.add_var(6); c.add_obj(lambda q: q.a * x[q.i] ** 2, d)
    aos = words_of(c)
    assert create(aos)[0] == 0
    k = 6 + 6                                        # header + (kind nitr itr_kind start databuf stride) -> nfields
    assert aos[k] == 2
    bad = list(aos); bad[k + 2] = 9                  # type of field 0
    assert create(bad)[0] == 3 and "field type" in create(bad)[1]
    bad = list(aos); bad[k + 1] = 12                 # byte offset of field 0: 12 + 8 > stride 16
    assert create(bad)[0] == 3 and "outside the iterator element" in create(bad)[1]
    # a DATA_SELF node inside an AoS pattern (tag 2): patch the first DATA_FIELD node (tag 3) of the tree
    nn = k + 1 + 2 * 2 + 4 + 1                       # fields, o0 o1 o2 base, nidx(=0)
    nnodes = aos[nn]
    nodes = nn + 1
    j = next(q for q in range(nnodes) if aos[nodes + 4 * q] == 3)
    bad = list(aos); bad[nodes + 4 * j] = 2
    assert create(bad)[0] == 3 and "DATA_SELF" in create(bad)[1]


def test_duplicate_free_hessian_closed_form_counts():
    """The fused form's unique count is a closed form of the pattern shifts: it must equal the number of distinct coordinates of
    the oracle's structure on random shift-indexed models; a distance whose columns have a gap switches the fused form off."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from edge_models import tile_fuzz
    for seed in range(6):
        core = tile_fuzz(seed)
        r, c = Oracle.from_core(core).hess_structure()
        info = E.Plan(core).tile_info()
        assert info["fused"] and info["nnzh_unique"] == len(set(zip(r.tolist(), c.tolist())))
    assert not E.Plan(tile_fuzz(99)).tile_info()["fused"]
    for core in (M.goddard_rocket(20), M.ac_power(M.synthetic_power_data(30, 41, 6, seed=5)), M.luksan_vlcek_aug(9, 3)):
        assert not E.Plan(core).tile_info()["fused"]        # fixed-index variable / data-indexed / product iterators


def test_sweep_gradient_is_planned_for_a_sole_shift_indexed_objective():
    """exb_eval writes g from its own sweep (P::EGRAD) exactly when the model has ONE objective pattern with gradient slots and
    that pattern is shift-indexed; several objective patterns, or indices from data / at fixed positions, keep the gradient launch."""
    import examodels_jl_b200 as E
    from examodels_jl_b200 import models as M
    from edge_models import EDGE, shared_targets
    assert "EGRAD = true" in E.Plan(M.luksan_vlcek(50)).source()
    assert "EGRAD = true" in E.Plan(EDGE["only_objective"]()).source()
    for core in (EDGE["mixed_gradient"](), shared_targets(100, 10), M.pattern_family(100, 32), M.goddard_rocket(10)):
        assert "EGRAD = true" not in E.Plan(core).source()


def test_hprod_contributions_are_grouped_per_distinct_variable():
    """P::hp (the point's Hessian-vector contributions): one entry per DISTINCT index expression -- LV constraint 3 (x[i], x[i+1],
    x[i+2]) for 6 slots, LV objective 2 for 3 slots -- with no run-time index compare where the shifts prove the indices different;
    data-indexed endpoints that may coincide (self loops) keep the compare so that the entry stays a diagonal one."""
    import re
    import examodels_jl_b200 as E
    from examodels_jl_b200 import models as M
    from edge_models import EDGE
    src = E.Plan(M.luksan_vlcek(50)).source()
    assert sorted(int(v) for v in re.findall(r"static constexpr int NT2 = (\d+);", src)) == [2, 3]
    hp = [src[m.start():src.index("static constexpr int PPT0", m.start())] for m in re.finditer(r"void hp\(", src)]
    assert len(hp) == 2 and all("if (idx[" not in h for h in hp)
    assert sum(h.count("__ldg(v +") for h in hp) == 5          # one v load per distinct variable of a point
    loops = E.Plan(EDGE["self_loops"]()).source()
    hp = [loops[m.start():loops.index("static constexpr int PPT0", m.start())] for m in re.finditer(r"void hp\(", loops)]
    assert any("if (idx[" in h for h in hp)
    rocket = E.Plan(M.goddard_rocket(10)).source()
    nt = [int(v) for v in re.findall(r"static constexpr int NT2 = (\d+);", rocket)]
    assert max(nt) <= 9            # h, v, m, T at two steps + the step variable: 9 scattered adds for the 47-slot dynamics pattern
