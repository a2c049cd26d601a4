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
