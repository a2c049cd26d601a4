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
