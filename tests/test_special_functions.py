"""SpecialFunctions-extension operators (/root/reference/ext/functionlist.jl): the reference's values come from
SpecialFunctions.jl / openspecfun (third party, not vendored), so both restatements -- the oracle's extended-precision one
(oracle/exa_oracle.cpp) and the device path's double-precision one (examodels.jl_b200/csrc/exb_special.h, compiled here for
the host) -- are pinned against mpmath values stored in tests/golden/special_functions.json."""
import ctypes as C
import json
import math
import os
import subprocess

import numpy as np
import pytest

import examodels_jl_b200.graph as G
from oracle.oracle_api import bi, uni

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "special_functions.json")))


def _rel(got, ref, name, x, tol, floor=0.0):
    err = abs(got - ref) / max(abs(ref), floor, 1e-300)
    assert err <= tol, f"{name}({x}): got {got!r} ref {ref!r} rel err {err:.2e}"


def _floor(name, x):
    """Oscillating functions are compared against their local amplitude near a zero crossing."""
    if name.startswith("airy") and x < 0:
        return abs(x) ** (0.25 if name.endswith("prime") else -0.25) / math.sqrt(math.pi)
    if name.startswith("bessel"):
        return 0.3 / math.sqrt(max(abs(x), 1.0))
    if name in ("digamma",):
        return 1.0
    return 0.0


@pytest.mark.parametrize("name", list(GOLD["univariate"]))
def test_oracle_special_values(name):
    op = G.OP1_CODE[name]
    for x, ref in GOLD["univariate"][name]:
        _rel(uni(op, x)[0], ref, name, x, 2e-13, _floor(name, x))


def test_oracle_polygamma_derivative_entries():
    for x, ref in GOLD["polygamma2"]:
        _rel(uni(G.OP1_CODE["digamma"], x)[2], ref, "polygamma2", x, 1e-13)
        _rel(uni(G.OP1_CODE["trigamma"], x)[1], ref, "polygamma2", x, 1e-13)
    for x, ref in GOLD["polygamma3"]:
        _rel(uni(G.OP1_CODE["trigamma"], x)[2], ref, "polygamma3", x, 1e-13)


@pytest.mark.parametrize("name", list(GOLD["bivariate"]))
def test_oracle_special_bivariate_values(name):
    for a, b, ref in GOLD["bivariate"][name]:
        _rel(bi(G.OP2_CODE[name], a, b)[0], ref, name, (a, b), 1e-13, 1e-3 if name == "logbeta" else 0.0)


@pytest.fixture(scope="module")
def host_special(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("sf") / "special_host.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests", "special_host.cpp")])
    L = C.CDLL(so)
    for n in ("digamma", "trigamma", "polygamma2", "polygamma3", "invdigamma", "dawson", "erfi"):
        getattr(L, "sf_" + n).restype = C.c_double
        getattr(L, "sf_" + n).argtypes = [C.c_double]
    L.sf_airy.restype = C.c_double
    L.sf_airy.argtypes = [C.c_double, C.c_int]
    return L


def test_device_special_algorithms_on_host(host_special):
    """exb_special.h (the hand-written part of the device path) compiled for the host: same pins as the oracle."""
    L = host_special
    f = {"digamma": L.sf_digamma, "trigamma": L.sf_trigamma, "invdigamma": L.sf_invdigamma, "dawson": L.sf_dawson, "erfi": L.sf_erfi,
         "airyai": lambda x: L.sf_airy(x, 0), "airyaiprime": lambda x: L.sf_airy(x, 1), "airybi": lambda x: L.sf_airy(x, 2),
         "airybiprime": lambda x: L.sf_airy(x, 3)}
    for name, fn in f.items():
        for x, ref in GOLD["univariate"][name]:
            _rel(fn(x), ref, name, x, 3e-13, _floor(name, x))
    for x, ref in GOLD["polygamma2"]:
        _rel(L.sf_polygamma2(x), ref, "polygamma2", x, 1e-13)
    for x, ref in GOLD["polygamma3"]:
        _rel(L.sf_polygamma3(x), ref, "polygamma3", x, 1e-13)


def test_device_and_oracle_special_agree_on_a_sweep(host_special):
    """Dense sweep: the two independent restatements agree far inside the 1e-10 parity tolerance."""
    L = host_special
    xs = np.linspace(-24.0, 24.0, 1921)
    for w, name in enumerate(("airyai", "airyaiprime", "airybi", "airybiprime")):
        op = G.OP1_CODE[name]
        for x in xs:
            _rel(L.sf_airy(float(x), w), uni(op, float(x))[0], name, x, 1e-12, _floor(name, x))
    for name, fn in (("dawson", L.sf_dawson), ("erfi", L.sf_erfi), ("digamma", L.sf_digamma), ("trigamma", L.sf_trigamma)):
        op = G.OP1_CODE[name]
        for x in np.linspace(-9.37, 9.4, 400):
            _rel(fn(float(x)), uni(op, float(x))[0], name, x, 1e-12, 1.0 if name == "digamma" else 0.0)


def test_special_ops_fold_constants_like_the_reference():
    """A Real argument is evaluated eagerly (register.jl:70): erf(0.5) is a number, not a node."""
    assert abs(G.erf(0.5) - math.erf(0.5)) < 1e-15 and abs(G.beta(2.0, 3.0) - 1.0 / 12.0) < 1e-15
    assert isinstance(G.erf(G.Var(1)), G.Node1) and isinstance(G.beta(G.Var(1), 2.0), G.Node2)


def _sweep_inputs(core):
    meta = core.meta()
    x = 0.004 * np.random.default_rng(4).uniform(-1.0, 1.0, meta["nvar"])
    y = np.random.default_rng(5).uniform(0.5, 1.5, meta["ncon"])
    return x, y


def test_oracle_special_sweep_vs_mpmath_derivatives():
    """f, f', f'' of every SpecialFunctions operator over its sweep: the oracle's cons / Jacobian / Hessian of `op(x + a)`
    against mpmath's numerical differentiation of the function itself (an independent check of the derivative FORMULAS)."""
    import importlib.util
    import mpmath as mp
    from edge_models import special_sweep, special_sweep_args
    from oracle.oracle_api import Oracle
    spec = importlib.util.spec_from_file_location("make_special_golden", os.path.join(ROOT, "tests", "golden", "make_special_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    F1 = gen.F1                                          # same mpmath definitions as the golden generator
    mp.mp.dps = 30
    n = 33
    core = special_sweep(n)
    o = Oracle.from_core(core)
    x, y = _sweep_inputs(core)
    c, j, h = o.cons(x), o.jac_coord(x), o.hess_coord(x, y, 0.0)
    for kp, name in enumerate(G.SPECIAL_UNIVARIATE):
        arg = x + special_sweep_args(name, n)
        for k in range(n):
            r = kp * n + k
            t = mp.mpf(float(arg[k]))
            f0, f1, f2 = (float(mp.diff(F1[name], t, q)) for q in (0, 1, 2))
            scale = max(abs(f0), abs(f1), abs(f2), 1e-300)
            ok = abs(c[r] - f0) <= 2e-9 * scale and abs(j[r] - f1) <= 2e-9 * scale and abs(h[r] - y[r] * f2) <= 8e-8 * scale
            assert ok, (name, float(arg[k]), (c[r], f0), (j[r], f1), (h[r] / y[r], f2))


@pytest.mark.gpu
def test_gpu_special_sweep_matches_oracle(exa):
    """Device special functions (libdevice + csrc/exb_special.h) over each operator's whole sweep: cons / jac / hess of
    `op(x + a)` against the oracle, POINT BY POINT relative to that point's own (f, f', f'') scale."""
    import torch
    from edge_models import special_sweep, special_sweep_args
    from oracle.oracle_api import Oracle
    n = 257
    core = special_sweep(n)
    ora, m = Oracle.from_core(core), exa.ExaModel(core)
    x, y = _sweep_inputs(core)
    dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    got = [m.cons_nln(dx, m.new(m.ncon)).cpu().numpy(), m.jac_coord(dx, m.new(m.nnzj)).cpu().numpy(),
           m.hess_coord(dx, dy, m.new(m.nnzh), obj_weight=0.0).cpu().numpy()[: m.ncon]]
    ref = [ora.cons(x), ora.jac_coord(x), ora.hess_coord(x, y, 0.0)[: ora.ncon]]
    scale = np.maximum.reduce([np.abs(r) for r in ref]) + 1e-300
    for g_, r_, what in zip(got, ref, ("cons", "jac", "hess")):
        err = np.abs(g_ - r_) / scale
        k = int(np.argmax(err))
        name = G.SPECIAL_UNIVARIATE[k // n]
        assert err[k] <= 1e-10, (f"{name} {what}: point {k % n} (arg {special_sweep_args(name, n)[k % n]:.4f}) got {g_[k]!r} "
                                 f"ref {r_[k]!r} rel err {err[k]:.2e}")
