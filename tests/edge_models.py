"""Edge-case models shared by the CPU (oracle) and GPU (parity) tests."""
import numpy as np

import examodels_jl_b200 as E
from examodels_jl_b200.graph import cos, exp, sin


def only_objective():
    c = E.ExaCore(); x = c.add_var(9, start=np.linspace(0.1, 0.9, 9))
    c.add_obj(lambda i: (x[i] - x[i + 1]) ** 2 * exp(x[i]), range(1, 9))
    return c


def only_constraints():
    c = E.ExaCore(); x = c.add_var(9, start=np.linspace(0.1, 0.9, 9))
    c.add_con(lambda i: sin(x[i]) * x[i + 1], range(1, 9))
    return c


def empty_patterns():
    """Zero-length iterators mixed with non-empty ones (a range of length 0 and an empty array)."""
    c = E.ExaCore(); x = c.add_var(6, start=np.linspace(0.2, 0.7, 6))
    c.add_con(lambda i: x[i] ** 3, range(1, 1))
    c.add_obj(lambda i: x[i] * x[i + 1], range(1, 6))
    c.add_con(lambda d: d.a * x[d.i] ** 2, np.zeros(0, dtype=np.dtype([("i", "i8"), ("a", "f8")])))
    c.add_con(lambda i: x[i] * exp(x[i + 1]), range(2, 5))
    c.add_obj(lambda i: x[i] ** 2, range(3, 3))
    return c


def single_points_and_constants():
    """One-row patterns, a fixed-index variable shared by every point, a constant body and Null rows."""
    c = E.ExaCore(); x = c.add_var(7, start=np.linspace(0.3, 0.9, 7)); t = c.add_var(1, start=0.5)
    c.add_obj(x[7] * t[1])
    c.add_con(x[1] - 1.0)
    c.add_con(lambda i: t[1] * (x[i] + x[i + 1]) ** 2, range(1, 7))
    c.add_con(lambda i: 3.5, range(1, 4))
    g = c.add_con(dims=(4,))
    c.add_con_aug(lambda i: g[i] + x[i] * x[i + 3], range(1, 5))
    return c


def self_loops():
    """AC-OPF-like cross terms whose two variable indices coincide numerically for some points: the slots stay
    distinct (dedupe is symbolic) and the cross slot gets 2*adj (src/hessian.jl:261-266)."""
    rng = np.random.default_rng(8)
    n = 40
    d = np.zeros(n, dtype=np.dtype([("f", "i8"), ("t", "i8"), ("g", "f8")]))
    d["f"] = rng.integers(1, 11, n); d["t"] = rng.integers(1, 11, n)
    d["t"][::3] = d["f"][::3]                      # every third branch is a self loop
    d["g"] = rng.uniform(0.5, 2.0, n)
    c = E.ExaCore(); v = c.add_var(10, start=rng.uniform(0.9, 1.1, 10)); a = c.add_var(10, start=rng.uniform(-0.2, 0.2, 10))
    c.add_con(lambda b: b.g * (v[b.f] * v[b.t] * cos(a[b.f] - a[b.t])), d)
    c.add_obj(lambda b: (v[b.f] - v[b.t]) ** 2 + v[b.f] * v[b.t], d)
    return c


def field_types():
    """int32 / float32 / int64 / float64 fields, an Int field used as a VALUE beyond 2^31, and a nested struct."""
    rng = np.random.default_rng(9)
    n = 33
    inner = np.dtype([("k", "i4"), ("w", "f4")])
    dt = np.dtype([("i", "i8"), ("big", "i8"), ("p", inner), ("a", "f8")])
    d = np.zeros(n, dtype=dt)
    d["i"] = np.arange(1, n + 1)
    d["big"] = 3_000_000_000 + np.arange(n)
    d["p"]["k"] = rng.integers(1, n + 1, n)
    d["p"]["w"] = rng.uniform(0.5, 1.5, n).astype(np.float32)
    d["a"] = rng.uniform(0.5, 1.5, n)
    c = E.ExaCore(); x = c.add_var(n + 1, start=rng.uniform(0.5, 1.0, n + 1))
    c.add_con(lambda q: q.p.w * x[q.i] * x[q.p.k] + q.big * 1e-9 * x[q.i + 1] ** 2, d)
    c.add_obj(lambda q: q.a * sin(x[q.p.k]) + q.i * x[q.i], d)
    return c


def mixed_gradient():
    """Objective patterns of every gradient flavour in one model: shift-indexed over ranges (owner-computes kernel,
    two of them overlapping on the same variables, shifts up to +-3), a strided index x[2i], a data-indexed one and a
    body too heavy to re-evaluate per slot (all three: slot + segmented-sum path), plus variables no objective touches."""
    rng = np.random.default_rng(10)
    n = 41
    c = E.ExaCore(); x = c.add_var(n, start=rng.uniform(0.4, 0.9, n)); z = c.add_var(5, start=rng.uniform(0.4, 0.9, 5))
    c.add_obj(lambda i: (x[i - 3] - x[i + 3]) ** 2 * x[i], range(4, n - 6))
    c.add_obj(lambda i: x[i] * x[i + 1] * x[i + 1 + 0] + 0.5 * x[1 + i], range(2, n - 1))
    c.add_obj(lambda i: x[2 * i] ** 3, range(1, n // 2))
    d = np.zeros(17, dtype=np.dtype([("i", "i8"), ("a", "f8")]))
    d["i"] = rng.integers(1, n + 1, 17); d["a"] = rng.uniform(0.5, 1.5, 17)
    c.add_obj(lambda q: q.a * sin(x[q.i]) * x[q.i], d)
    c.add_obj(lambda i: exp(sin(x[i]) * cos(x[i + 1])) * exp(x[i + 2]) * sin(x[i + 3] * x[i]) * cos(exp(x[i + 1] - x[i + 2]))
              * sin(cos(x[i]) + x[i + 3]) * exp(-x[i + 1] * x[i + 2]) * cos(sin(x[i + 3]) - x[i]) * sin(exp(x[i]) * 0.1)
              * cos(x[i + 1] * x[i + 3]) * exp(sin(x[i + 2])), range(1, 9))
    c.add_con(lambda i: x[i] * z[1] + x[i + 1], range(1, 6))
    return c


SPECIAL_SWEEPS = {   # argument sweeps inside each SpecialFunctions operator's domain, away from poles (ext/functionlist.jl)
    "erf": (-4.0, 4.0), "erfc": (-3.0, 5.0), "erfi": (-2.5, 2.5), "erfcx": (-2.0, 30.0), "digamma": (-3.63, 14.0),
    "trigamma": (-3.63, 14.0), "invdigamma": (-4.0, 3.0), "gamma": (-3.63, 6.0), "airyai": (-12.0, 4.0), "airybi": (-12.0, 4.0),
    "airyaiprime": (-12.0, 4.0), "airybiprime": (-12.0, 4.0), "besselj0": (-15.0, 15.0), "bessely0": (0.3, 25.0),
    "besselj1": (-15.0, 15.0), "bessely1": (0.3, 25.0), "dawson": (-9.0, 9.0), "erfinv": (-0.98, 0.98), "erfcinv": (0.02, 1.98),
}


def special_sweep_args(name, n):
    lo, hi = SPECIAL_SWEEPS[name]
    a = np.linspace(lo, hi, n)
    if name in ("digamma", "trigamma", "gamma"):   # nudge points within 0.12 of a pole (non-positive integers) away
        near = (a < 0.5) & (np.abs(a - np.round(a)) < 0.12)
        a = np.where(near, np.round(a) + 0.37, a)
    return a


def special_sweep(n=257):
    """One constraint pattern `op(x[i] + a_i)` per SpecialFunctions operator, a_i sweeping the operator's domain (x starts
    at 0).  Pattern k owns rows / Jacobian slots / Hessian slots [k n, (k + 1) n)."""
    from examodels_jl_b200 import graph as G
    c = E.ExaCore(); x = c.add_var(n, start=np.zeros(n))
    for name in G.SPECIAL_UNIVARIATE:
        d = np.zeros(n, dtype=np.dtype([("i", "i8"), ("a", "f8")]))
        d["i"] = np.arange(1, n + 1); d["a"] = special_sweep_args(name, n)
        c.add_con(lambda p, nm=name: G._op1(nm, x[p.i] + p.a), d)
    c.add_obj(lambda i: x[i] ** 2, range(1, n + 1))
    return c


def shared_targets(n=20000, k=700):
    """Targets that thousands of slots land on: two variables at FIXED indices shared by every point (a step length: gradient,
    Jacobian-column and Hessian row/column runs of n slots) and rows that collect n / k / 40 augmentation terms.  The one-thread-
    per-target sums of the reference (ext:482-511,691-697) would be serial over them: runs above 64 slots are summed by a warp,
    runs of 8192 or more by chunks (csrc/exb_fixed.cu); the fused products aggregate their atomics per block."""
    rng = np.random.default_rng(13)
    c = E.ExaCore(); x = c.add_var(n + 1, start=rng.uniform(0.4, 0.9, n + 1)); t = c.add_var(2, start=[0.5, 0.8])
    c.add_obj(lambda i: t[1] * (x[i] - x[i + 1]) ** 2 + t[2] * x[i], range(1, n + 1))
    c.add_con(lambda i: x[i + 1] - x[i] - t[1] * sin(x[i]) * t[2], range(1, n + 1))
    g = c.add_con(dims=(3,))
    d = np.zeros(n + k + 40, dtype=np.dtype([("r", "i8"), ("i", "i8"), ("a", "f8")]))
    d["r"] = np.concatenate([np.full(n, 1), np.full(k, 2), np.full(40, 3)])
    d["i"] = rng.integers(1, n + 2, len(d)); d["a"] = rng.uniform(0.5, 1.5, len(d))
    rng.shuffle(d)
    c.add_con_aug(lambda q: g[q.r] + q.a * x[q.i] * t[2], d)
    return c


EDGE = {"mixed_gradient": mixed_gradient, "only_objective": only_objective, "only_constraints": only_constraints, "empty_patterns": empty_patterns,
        "single_points_and_constants": single_points_and_constants, "self_loops": self_loops, "field_types": field_types}


def tile_fuzz(seed):
    """Random shift-indexed models for the column-tile kernels (duplicate-free Hessian, tile gradient): 2-4 patterns over
    different sub-ranges (so the per-distance column intervals start / end at different places, sometimes with gaps -> the
    sorted-gather fallback), shifts in [-3, 3], bodies mixing products, powers, sin / exp, parameters-free."""
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(150, 900))
    c = E.ExaCore(); x = c.add_var(n, start=rng.uniform(0.4, 0.9, n))
    from examodels_jl_b200.graph import sin, exp, cos
    if seed == 99:   # two patterns over disjoint ranges: the columns of distance 1 have a GAP -> no closed form, sorted gather
        c.add_con(lambda i: x[i] * x[i + 1] + sin(x[i]), range(1, 50))
        c.add_obj(lambda i: x[i] * x[i + 1] ** 2, range(100, 140))
        return c
    npat = int(rng.integers(2, 5))
    for k in range(npat):
        sh = sorted(set(int(v) for v in rng.integers(-3, 4, size=int(rng.integers(2, 4)))))
        lo = 1 - min(sh) + int(rng.integers(0, 40)); hi = n - max(sh) - int(rng.integers(0, 40))
        if rng.random() < 0.3:                       # a short pattern somewhere in the middle
            lo = int(rng.integers(lo, (lo + hi) // 2)); hi = int(rng.integers(lo + 5, hi))
        kind = int(rng.integers(0, 3))
        a, b = sh[0], sh[-1]
        m = sh[len(sh) // 2]

        def body(i, a=a, b=b, m=m, kind=kind):
            t = x[i + a] * x[i + b] ** 2 + sin(x[i + m] - x[i + a])
            if kind == 1:
                t = t * exp(x[i + b] * 0.3) + x[i + m] ** 3
            if kind == 2:
                t = t + cos(x[i + a] * x[i + m]) * x[i + b]
            return t
        if k % 2 == 0:
            c.add_con(body, range(lo, hi + 1))
        else:
            c.add_obj(body, range(lo, hi + 1))
    return c
