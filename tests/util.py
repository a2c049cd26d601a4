"""Shared helpers for the parity tests: seeded inputs and the comparison rule."""
import numpy as np

RTOL = 1e-10  # BASELINE.json north_star: FP64 derivative values within 1e-10 relative


def inputs(core, seed=0):
    """x = x0 + 0.01 u, u ~ U(-1,1) (seed); y ~ N(0,1) (seed+1)."""
    meta = core.meta()
    rng = np.random.default_rng(seed)
    x = meta["x0"] + 0.01 * rng.uniform(-1.0, 1.0, meta["nvar"])
    y = np.random.default_rng(seed + 1).standard_normal(meta["ncon"])
    return np.ascontiguousarray(x), np.ascontiguousarray(y)


def assert_close(got, ref, what="", rtol=RTOL):
    """|got - ref| <= rtol * max(|ref_i|, ||ref||_inf): entrywise relative where the entry carries the
    vector's scale, norm-wise for entries that are (near) cancellations of larger terms."""
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    if ref.size == 0:
        return
    assert np.array_equal(np.isnan(got), np.isnan(ref)), f"{what}: NaN pattern differs"
    m = ~np.isnan(ref)
    scale = float(np.max(np.abs(ref[m]))) if m.any() else 0.0
    bound = rtol * np.maximum(np.abs(ref[m]), scale)
    err = np.abs(got[m] - ref[m])
    bad = err > bound
    if bad.any():
        k = int(np.argmax(err / np.maximum(bound, 1e-300)))
        raise AssertionError(f"{what}: {int(bad.sum())} of {ref.size} entries off; worst idx {k}: got {got[m][k]!r} "
                             f"ref {ref[m][k]!r} err {err[k]:.3e} bound {bound[k]:.3e}")


# ---- replay of the reference's printed Ipopt runs (tests/golden/ipopt_logs.json) --------------------------------------------
def newton_kkt_replay(cb, x0, niter):
    """What Ipopt does on an equality-constrained problem without bounds when no regularisation and no step cut is needed
    (the logs say so: lg(rg) '-', alpha 1, ls 1): objective scaling s = min(1, 100 / max|grad f(x0)|), least-squares initial
    multipliers, then full Newton steps on the KKT system with the exact Lagrangian Hessian.  `cb` provides obj(x),
    grad(x), cons(x), jac(x) (dense) and hess(x, lam, sigma) (dense, symmetric).  Returns the rows Ipopt prints."""
    x = np.array(x0, dtype=np.float64)
    g, J = cb["grad"](x), cb["jac"](x)
    s = min(1.0, 100.0 / np.abs(g).max())
    lam = -np.linalg.solve(J @ J.T, J @ g)
    rows = [dict(objective=cb["obj"](x), inf_pr=np.abs(cb["cons"](x)).max(), inf_du=s * np.abs(g + J.T @ lam).max(), d_norm=0.0)]
    n, m = x.size, lam.size
    for _ in range(niter):
        g, J, c, H = cb["grad"](x), cb["jac"](x), cb["cons"](x), cb["hess"](x, lam, 1.0)
        K = np.block([[H, J.T], [J, np.zeros((m, m))]])
        d = np.linalg.solve(K, -np.concatenate([g + J.T @ lam, c]))
        x, lam = x + d[:n], lam + d[n:]
        g, J = cb["grad"](x), cb["jac"](x)
        rows.append(dict(objective=cb["obj"](x), inf_pr=np.abs(cb["cons"](x)).max(), inf_du=s * np.abs(g + J.T @ lam).max(),
                         d_norm=np.abs(d[:n]).max()))
    final = dict(scaled_objective=s * cb["obj"](x), unscaled_objective=cb["obj"](x), constraint_violation=np.abs(cb["cons"](x)).max(),
                 unscaled_dual_infeasibility=np.abs(g + J.T @ lam).max())
    return rows, final, x, lam


def assert_matches_printed(value, printed, what, noise=0.0):
    """`printed` is a number as Ipopt printed it (e.g. '1.0953147e+03'): equal up to the last printed digit (+ `noise`, the
    absolute rounding noise of evaluating that quantity -- it matters only for the converged iterate's residuals)."""
    mant, exp = printed.lower().split("e")
    digits = len(mant.replace("-", "").replace(".", "")) - 1
    ulp = 10.0 ** (int(exp) - digits)
    assert abs(value - float(printed)) <= 0.6 * ulp + noise, f"{what}: {value!r} vs printed {printed}"


def check_ipopt_run(run, rows, final):
    assert len(rows) == len(run["iterations"])
    for k, (ours, ref) in enumerate(zip(rows, run["iterations"])):
        assert ref["lg_rg"] == "-" and (k == 0 or (ref["alpha_pr"] == "1.00e+00" and ref["alpha_du"] == "1.00e+00" and ref["ls"] == 1))
        for col, noise in (("objective", 0.0), ("inf_pr", 1e-14), ("inf_du", 2e-13), ("d_norm", 0.0)):
            assert_matches_printed(ours[col], ref[col], f"theta {run['theta']} iter {k} {col}", noise)
    f = run["final"]
    assert abs(final["scaled_objective"] - float(f["objective"]["scaled"])) <= 1e-12 * abs(float(f["objective"]["scaled"]))
    assert abs(final["unscaled_objective"] - float(f["objective"]["unscaled"])) <= 1e-12 * abs(float(f["objective"]["unscaled"]))
    assert abs(final["constraint_violation"] - float(f["constraint_violation"]["unscaled"])) <= 2e-2 * float(f["constraint_violation"]["unscaled"])
    assert abs(final["unscaled_dual_infeasibility"] - float(f["dual_infeasibility"]["unscaled"])) <= 2e-2 * float(f["dual_infeasibility"]["unscaled"])
