"""Model builders for the configurations BASELINE.json names.

Each builder writes the model the way the reference's own sources write it, so the
tree shapes (and therefore the sparsity layout) match:

* `luksan_vlcek`        — docs/src/guide.jl:32-50 and benchmark/runbenchmark.jl:163-169
                          (`order="bench"`: constraints first, as the benchmark does)
* `luksan_vlcek_aug`    — test/NLPTest/luksan.jl:17-26 (base con1 + augmentation con2,
                          2-D variable block x[N, M], product iterators)
* `ac_power`            — test/NLPTest/power.jl:112-213 over a synthetic network
* `goddard_rocket`      — README.md:19-26 velocity pattern + the public COPS 3.0 statement
                          (parity unpinned in the reference tree)
* `pattern_family`      — 32 structurally distinct patterns
"""
from __future__ import annotations

import numpy as np

from . import graph as G
from .graph import cos, exp, sin, sqrt, log, tanh, atan, cosh  # noqa: F401
from .nlp import ExaCore, product


# ---------------------------------------------------------------------------
# Lukšan–Vlček
# ---------------------------------------------------------------------------
def lv_x0(N):
    i = np.arange(1, N + 1)
    return np.where(i % 2 == 1, -1.2, 1.0)


def luksan_vlcek(N, order="bench"):
    c = ExaCore()
    x = c.add_var(N, start=lv_x0(N))

    def con(i):
        return (3 * x[i + 1] ** 3 + 2 * x[i + 2] - 5
                + sin(x[i + 1] - x[i + 2]) * sin(x[i + 1] + x[i + 2]) + 4 * x[i + 1]
                - x[i] * exp(x[i] - x[i + 1]) - 3)

    def obj(i):
        return 100 * (x[i - 1] ** 2 - x[i]) ** 2 + (x[i - 1] - 1) ** 2

    if order == "bench":  # benchmark/runbenchmark.jl:166-167
        c.add_con(con, range(1, N - 1))
        c.add_obj(obj, range(2, N + 1))
    else:  # docs/src/guide.jl:40-50
        c.add_obj(obj, range(2, N + 1))
        c.add_con(con, range(1, N - 1))
    return c


def luksan_vlcek_param(N, theta=(100.0, 1.0)):
    """docs/src/parameters.md:20-90 of the reference: LV with the penalty coefficient and the offset of the objective as
    parameters, `θ[1] * (x[i-1]^2 - x[i])^2 + (x[i-1] - θ[2])^2` (objective added first, as there)."""
    c = ExaCore()
    th = c.add_par(list(theta))
    x = c.add_var(N, start=lv_x0(N))
    c.add_obj(lambda i: th[1] * (x[i - 1] ** 2 - x[i]) ** 2 + (x[i - 1] - th[2]) ** 2, range(2, N + 1))
    c.add_con(lambda i: (3 * x[i + 1] ** 3 + 2 * x[i + 2] - 5
                         + sin(x[i + 1] - x[i + 2]) * sin(x[i + 1] + x[i + 2]) + 4 * x[i + 1]
                         - x[i] * exp(x[i] - x[i + 1]) - 3), range(1, N - 1))
    return c


def luksan_vlcek_aug(N, M=1):
    """test/NLPTest/luksan.jl:17-26."""
    c = ExaCore()
    i0 = np.arange(1, N + 1)
    start = np.repeat(np.where(i0 % 2 == 1, -1.2, 1.0)[:, None], M, axis=1)
    x = c.add_var(N, M, start=start)
    itr = product(range(1, N - 1), range(1, M + 1))

    def con1(d):
        i, j = d
        return 3 * x[i + 1, j] ** 3 + 2 * x[i + 2, j] - 5

    def con2(d):
        i, j = d
        return ((i, j),
                sin(x[i + 1, j] - x[i + 2, j]) * sin(x[i + 1, j] + x[i + 2, j]) + 4 * x[i + 1, j]
                - x[i, j] * exp(x[i, j] - x[i + 1, j]) - 3)

    def obj(d):
        i, j = d
        return 100 * (x[i - 1, j] ** 2 - x[i, j]) ** 2 + (x[i - 1, j] - 1) ** 2

    s = c.add_con(con1, itr)
    c.add_con_aug(s, con2, itr)
    c.add_obj(obj, product(range(2, N + 1), range(1, M + 1)))
    return c


# ---------------------------------------------------------------------------
# AC-OPF-shaped pattern set (test/NLPTest/power.jl:112-213)
# ---------------------------------------------------------------------------
def synthetic_power_data(nbus=10_000, nbranch=14_000, ngen=2_500, seed=2):
    """Random radial-plus-chords topology."""
    rng = np.random.default_rng(seed)
    assert nbranch >= nbus - 1
    f = np.empty(nbranch, dtype=np.int64)
    t = np.empty(nbranch, dtype=np.int64)
    # spanning tree: bus k (k>=2) hangs off a random earlier bus
    t[: nbus - 1] = np.arange(2, nbus + 1)
    f[: nbus - 1] = (rng.random(nbus - 1) * np.arange(1, nbus)).astype(np.int64) + 1
    extra = nbranch - (nbus - 1)
    a = rng.integers(1, nbus + 1, size=extra)
    b = rng.integers(1, nbus, size=extra)
    b = np.where(b >= a, b + 1, b)  # b != a
    f[nbus - 1:], t[nbus - 1:] = a, b
    coef = rng.uniform(-10.0, 10.0, size=(nbranch, 8))
    branch = np.zeros(nbranch, dtype=np.dtype([
        ("i", "i8"), ("j", "i8"), ("f_idx", "i8"), ("t_idx", "i8"), ("f_bus", "i8"), ("t_bus", "i8"),
        ("c1", "f8"), ("c2", "f8"), ("c3", "f8"), ("c4", "f8"), ("c5", "f8"), ("c6", "f8"),
        ("c7", "f8"), ("c8", "f8"), ("rate_a_sq", "f8")]))
    branch["i"] = np.arange(1, nbranch + 1)
    branch["j"] = 1
    branch["f_idx"] = np.arange(1, nbranch + 1)  # arcs: 1..nbranch "from", nbranch+1.. "to"
    branch["t_idx"] = np.arange(nbranch + 1, 2 * nbranch + 1)
    branch["f_bus"], branch["t_bus"] = f, t
    for k in range(8):
        branch[f"c{k + 1}"] = coef[:, k]
    rate_a = rng.uniform(1.0, 10.0, size=nbranch)
    branch["rate_a_sq"] = rate_a ** 2
    arc = np.zeros(2 * nbranch, dtype=np.dtype([("i", "i8"), ("rate_a", "f8"), ("bus", "i8")]))
    arc["i"] = np.arange(1, 2 * nbranch + 1)
    arc["rate_a"] = np.concatenate([rate_a, rate_a])
    arc["bus"] = np.concatenate([f, t])
    bus = np.zeros(nbus, dtype=np.dtype([("i", "i8"), ("pd", "f8"), ("gs", "f8"), ("qd", "f8"), ("bs", "f8")]))
    bus["i"] = np.arange(1, nbus + 1)
    bus["pd"], bus["qd"] = rng.uniform(0, 1, nbus), rng.uniform(0, 1, nbus)
    bus["gs"], bus["bs"] = rng.uniform(0, 0.1, nbus), rng.uniform(0, 0.1, nbus)
    gen = np.zeros(ngen, dtype=np.dtype([("i", "i8"), ("cost1", "f8"), ("cost2", "f8"), ("cost3", "f8"), ("bus", "i8")]))
    gen["i"] = np.arange(1, ngen + 1)
    gen["cost1"], gen["cost2"], gen["cost3"] = rng.uniform(0, 1, ngen), rng.uniform(1, 50, ngen), rng.uniform(0, 10, ngen)
    gen["bus"] = rng.integers(1, nbus + 1, size=ngen)
    return dict(bus=bus, gen=gen, arc=arc, branch=branch, ref_buses=np.array([1], dtype=np.int64),
                vmax=np.full(nbus, 1.1), vmin=np.full(nbus, 0.9),
                pmax=np.full(ngen, 5.0), pmin=np.zeros(ngen), qmax=np.full(ngen, 5.0), qmin=np.full(ngen, -5.0),
                rate_a=arc["rate_a"].copy(), angmax=np.full(nbranch, 0.5), angmin=np.full(nbranch, -0.5))


def ac_power(data):
    """The 15 patterns of test/NLPTest/power.jl:112-207, in the same add order."""
    w = ExaCore()
    nb, ng, na = len(data["bus"]), len(data["gen"]), len(data["arc"])
    va = w.add_var(nb)
    vm = w.add_var(nb, start=np.ones(nb), lvar=data["vmin"], uvar=data["vmax"])
    pg = w.add_var(ng, lvar=data["pmin"], uvar=data["pmax"])
    qg = w.add_var(ng, lvar=data["qmin"], uvar=data["qmax"])
    p = w.add_var(na, lvar=-data["rate_a"], uvar=data["rate_a"])
    q = w.add_var(na, lvar=-data["rate_a"], uvar=data["rate_a"])
    br = data["branch"]

    w.add_obj(lambda g: g.cost1 * pg[g.i] ** 2 + g.cost2 * pg[g.i] + g.cost3, data["gen"])
    w.add_con(lambda i: va[i], data["ref_buses"])
    w.add_con(lambda b: p[b.f_idx] - b.c5 * vm[b.f_bus] ** 2
              - b.c3 * (vm[b.f_bus] * vm[b.t_bus] * cos(va[b.f_bus] - va[b.t_bus]))
              - b.c4 * (vm[b.f_bus] * vm[b.t_bus] * sin(va[b.f_bus] - va[b.t_bus])), br)
    w.add_con(lambda b: q[b.f_idx] + b.c6 * vm[b.f_bus] ** 2
              + b.c4 * (vm[b.f_bus] * vm[b.t_bus] * cos(va[b.f_bus] - va[b.t_bus]))
              - b.c3 * (vm[b.f_bus] * vm[b.t_bus] * sin(va[b.f_bus] - va[b.t_bus])), br)
    w.add_con(lambda b: p[b.t_idx] - b.c7 * vm[b.t_bus] ** 2
              - b.c1 * (vm[b.t_bus] * vm[b.f_bus] * cos(va[b.t_bus] - va[b.f_bus]))
              - b.c2 * (vm[b.t_bus] * vm[b.f_bus] * sin(va[b.t_bus] - va[b.f_bus])), br)
    w.add_con(lambda b: q[b.t_idx] + b.c8 * vm[b.t_bus] ** 2
              + b.c2 * (vm[b.t_bus] * vm[b.f_bus] * cos(va[b.t_bus] - va[b.f_bus]))
              - b.c1 * (vm[b.t_bus] * vm[b.f_bus] * sin(va[b.t_bus] - va[b.f_bus])), br)
    w.add_con(lambda b: va[b.f_bus] - va[b.t_bus], br, lcon=data["angmin"], ucon=data["angmax"])
    w.add_con(lambda b: p[b.f_idx] ** 2 + q[b.f_idx] ** 2 - b.rate_a_sq, br, lcon=-np.inf)
    w.add_con(lambda b: p[b.t_idx] ** 2 + q[b.t_idx] ** 2 - b.rate_a_sq, br, lcon=-np.inf)
    c9 = w.add_con(lambda b: b.pd + b.gs * vm[b.i] ** 2, data["bus"])
    c10 = w.add_con(lambda b: b.qd - b.bs * vm[b.i] ** 2, data["bus"])
    w.add_con_aug(c9, lambda a: (a.bus, p[a.i]), data["arc"])
    w.add_con_aug(c10, lambda a: (a.bus, q[a.i]), data["arc"])
    w.add_con_aug(c9, lambda g: (g.bus, -pg[g.i]), data["gen"])
    w.add_con_aug(c10, lambda g: (g.bus, -qg[g.i]), data["gen"])
    return w


# ---------------------------------------------------------------------------
# COPS Goddard rocket (README.md:19-26 for the velocity pattern; the rest from the
# public COPS 3.0 statement — not pinned by anything in the reference tree)
# ---------------------------------------------------------------------------
def goddard_rocket(nh):
    h_0, v_0, m_0, g_0 = 1.0, 0.0, 1.0, 1.0
    T_c, h_c, v_c, m_c = 3.5, 500.0, 620.0, 0.6
    c_ = 0.5 * np.sqrt(g_0 * h_0)
    m_f = m_c * m_0
    D_c = 0.5 * v_c * (m_0 / g_0)
    T_max = T_c * m_0 * g_0
    core = ExaCore(minimize=False)
    k = np.arange(0, nh + 1)
    h = core.add_var(range(0, nh + 1), start=np.ones(nh + 1), lvar=h_0)
    v = core.add_var(range(0, nh + 1), start=k / nh * (1.0 - k / nh), lvar=0.0)
    m = core.add_var(range(0, nh + 1), start=(m_f - m_0) * (k / nh) + m_0, lvar=m_f, uvar=m_0)
    tau = core.add_var(range(0, nh + 1), start=np.full(nh + 1, T_max / 2), lvar=0.0, uvar=T_max)
    dt = core.add_var(1, start=1.0 / nh, lvar=0.0)
    core.add_obj(h[nh])  # maximise final altitude (single-row pattern)

    def drag(i):
        return D_c * v[i] ** 2 * exp(-h_c * (h[i] - h_0) / h_0)

    def grav(i):
        return g_0 * (h_0 / h[i]) ** 2

    core.add_con(lambda i: -h[i] + h[i - 1] + 0.5 * dt[1] * (v[i] + v[i - 1]), range(1, nh + 1))
    core.add_con(lambda i: -v[i] + v[i - 1] + 0.5 * dt[1] * (
        (tau[i] - drag(i) - m[i] * grav(i)) / m[i]
        + (tau[i - 1] - drag(i - 1) - m[i - 1] * grav(i - 1)) / m[i - 1]), range(1, nh + 1))
    core.add_con(lambda i: -m[i] + m[i - 1] - 0.5 * dt[1] * (tau[i] + tau[i - 1]) / c_, range(1, nh + 1))
    # boundary conditions as single-row patterns
    core.add_con(h[0] - h_0)
    core.add_con(v[0] - v_0)
    core.add_con(m[0] - m_0)
    core.add_con(m[nh] - m_f)
    return core


# ---------------------------------------------------------------------------
# 32 structurally distinct patterns x n points (config 5)
# ---------------------------------------------------------------------------
_U = [sin, cos, exp, tanh, atan, cosh, lambda z: z ** 2, lambda z: z ** 3]


def pattern_family(n=1_000_000, npat=32, seed=3):
    """AoS iterator of (Int64 i, Float64 a, Float64 b); pattern k mixes ops/arity over
    x[i], x[i+1], x[i+2].  Even k are constraints, odd k objectives."""
    rng = np.random.default_rng(seed)
    c = ExaCore()
    x = c.add_var(n + 2, start=rng.uniform(0.5, 1.5, n + 2))
    dt = np.dtype([("i", "i8"), ("a", "f8"), ("b", "f8")])
    for k in range(npat):
        d = np.zeros(n, dtype=dt)
        d["i"] = np.arange(1, n + 1)
        d["a"] = rng.uniform(0.5, 2.0, n)
        d["b"] = rng.uniform(-1.0, 1.0, n)
        u1, u2 = _U[k % 8], _U[(k // 8 + 3 * k + 1) % 8]
        arity = 1 + (k % 3)

        def body(p, u1=u1, u2=u2, arity=arity, k=k):
            t = p.a * u1(x[p.i])
            if arity >= 2:
                t = t + u2(x[p.i] * x[p.i + 1]) * p.b
            if arity >= 3:
                t = t + x[p.i + 2] / (1 + x[p.i] ** 2) - p.b * x[p.i + 1] * x[p.i + 2]
            if k % 4 == 3:
                t = t * (x[p.i + 1] - p.a)
            return t

        if k % 2 == 0:
            c.add_con(body, d)
        else:
            c.add_obj(body, d)
    return c


# ---------------------------------------------------------------------------
# small coverage models (parameters; every registered operator)
# ---------------------------------------------------------------------------
def parametric(n=200, seed=4):
    """Parameters θ in objectives and constraints (add_par, src/nlp.jl:1177-1215; test/NLPTest/feature_test.jl:100-126)."""
    rng = np.random.default_rng(seed)
    c = ExaCore()
    x = c.add_var(n, start=rng.uniform(0.5, 1.5, n))
    th = c.add_par(n, value=rng.uniform(1.0, 2.0, n))
    w = c.add_par(range(2, 5), value=[10.0, 20.0, 30.0])
    c.add_obj(lambda i: th[i] * (x[i] - w[2]) ** 2 + exp(th[i] * x[i]), range(1, n + 1))
    c.add_con(lambda i: th[i] * x[i] * x[i + 1] - w[3] * sin(x[i] / th[i + 1]), range(1, n))
    c.add_con(lambda j: w[j] * x[1] + x[j] ** 3, range(2, 5))
    return c


def all_ops(n=64, part=0, seed=6):
    """One constraint pattern per registered univariate operator (src/functionlist.jl:6-60) applied to an
    argument inside its domain, and one per bivariate operator (:71-81) in its node-node, node-Real and
    Real-node forms: the operator coverage matrix of test/ADTest/ADTest.jl."""
    rng = np.random.default_rng(seed)
    c = ExaCore()
    x = c.add_var(n + 1, start=rng.uniform(0.3, 0.7, n + 1))
    d = np.zeros(n, dtype=np.dtype([("i", "i8"), ("a", "f8")]))
    d["i"] = np.arange(1, n + 1)
    d["a"] = rng.uniform(1.1, 1.9, n)
    big = {"acosh", "acoth"}                      # need |arg| > 1
    # parts keep each generated module small (compile time grows quickly with the pattern count); part 3 is the
    # SpecialFunctions extension (ext/functionlist.jl): 19 univariate + 2 bivariate operators
    nb = G.UNIVARIATE.index("erf")
    uni = G.UNIVARIATE[:26] if part == 0 else G.UNIVARIATE[26:nb] if part == 1 else G.UNIVARIATE[nb:] if part == 3 else []
    for name in uni:
        f = (lambda z, nm=name: G._op1(nm, z))
        if name in big:
            c.add_con(lambda p, f=f: f(p.a + x[p.i] * x[p.i + 1]) * x[p.i], d)
        else:
            c.add_con(lambda p, f=f: f(x[p.i] * x[p.i + 1]) * x[p.i] + p.a, d)
    for name in ([b for b in G.BIVARIATE if b not in G.SPECIAL_BIVARIATE] if part == 2 else G.SPECIAL_BIVARIATE if part == 3 else []):
        f = (lambda u, v, nm=name: G._op2(nm, u, v))
        c.add_con(lambda p, f=f: f(x[p.i] + 0.5, x[p.i + 1] * p.a) * x[p.i], d)       # node, node
        c.add_con(lambda p, f=f: f(x[p.i] * x[p.i + 1] + 0.5, p.a) * x[p.i + 1], d)   # node, Real
        c.add_con(lambda p, f=f: f(p.a, x[p.i] * x[p.i + 1] + 0.5) * x[p.i + 1], d)   # Real, node
    c.add_con(lambda p: (x[p.i] * p.a) ** 5 + G.pow_runtime(x[p.i + 1] + 1.0, p.i) / (1.0 + x[p.i] ** 2), d)  # Val and Int-data exponents
    c.add_obj(lambda p: sqrt(1.0 + x[p.i] ** 2) * log(p.a + x[p.i + 1]), d)
    return c
