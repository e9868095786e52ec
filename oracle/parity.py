"""TEST INFRASTRUCTURE ONLY: the parity protocol (DESIGN.md "Parity").

north_star asks for bit-exact grid-interval indices and Z / lPz within 1e-12 relative.  The reference's
quadratic formula cancels (tt_irt1_int32.c:150-156), so the reference ITSELF moves by up to ~1e-10 in a
few Z entries when only its BLAS summation order changes (re-measured by
tests/test_oracle.py).  The protocol therefore is, per entry (m, k):

    idx   : equal to the oracle's, except where the oracle reports q within `gap_tol` of a CDF node
    lPz   : |d| <= 1e-12 * max(1, |lPz|) + sum_k lsens_k * CFAC * eps * cumsum_k(cond)
    Z     : |d| <= 1e-12 * max(1, |Z|) + CFAC * eps * cumsum_k(cond)

cond is the oracle's first-order sensitivity of x_k to O(eps) perturbations of (cdf, p) including the
formula's cancellation factor; the cumulative sum carries an ill-conditioned early coordinate into the
later ones.  lsens_k = |d log p(x_k) / d x_k| of the interpolated conditional carries the same admitted
perturbation of x_k into the log-density: an entry whose Z is ill-conditioned has an equally ill-conditioned lPz
(the reference's own two BLAS builds show lPz differences of 0.4 * |dZ| on such entries, tests/devtools/lpz_outlier.py).
The reference's own OpenBLAS-vs-netlib spread sits below 2.5 * eps * cumsum(cond) on every
BASELINE shape (tests/test_oracle.py asserts that); CFAC = 3 is that bound plus a fifth (round 1 used 8; the fast GPU
path measures 0.08-0.15 of that, i.e. 0.2-0.4 of the present bound).
"""
import numpy as np

EPS = np.finfo(np.float64).eps
CFAC = 3.0
RTOL = 1e-12


def z_tolerance(Z_ref, cond):
    return RTOL * np.maximum(1.0, np.abs(Z_ref)) + CFAC * EPS * np.cumsum(cond, axis=1)


def lpz_tolerance(lPz_ref, cond, lsens):
    tol = RTOL * np.maximum(1.0, np.abs(lPz_ref))
    if lsens is not None:
        with np.errstate(invalid="ignore"):
            extra = np.nansum(np.where(np.isfinite(lsens), lsens, 0.0) * CFAC * EPS * np.cumsum(cond, axis=1), axis=1)
        tol = tol + extra
    return tol


def compare(Z, lPz, idx, Z_ref, lPz_ref, idx_ref, cond, gap, lsens=None, gap_tol=1e-13):
    """Returns a dict of parity statistics and a list of failure strings (empty = parity holds)."""
    fails = []
    stats = {}
    if idx is not None:
        flips = idx != idx_ref
        stats["idx_flips"] = int(flips.sum())
        hard = flips & (gap > gap_tol)
        stats["idx_flips_not_at_node"] = int(hard.sum())
        if hard.any():
            fails.append("%d interval indices differ away from CDF nodes" % int(hard.sum()))
        ok_rows = ~flips.any(axis=1)
    else:
        ok_rows = np.ones(Z.shape[0], dtype=bool)
    dz = np.abs(Z - Z_ref)
    tol = z_tolerance(Z_ref, cond)
    bad = (dz > tol) | ~np.isfinite(Z)
    bad &= ok_rows[:, None]
    bad &= np.isfinite(Z_ref)
    stats["z_max_abs"] = float(np.nanmax(dz[ok_rows])) if ok_rows.any() else 0.0
    stats["z_frac_gt_1e-12"] = float((dz[ok_rows] > RTOL * np.maximum(1.0, np.abs(Z_ref[ok_rows]))).mean()) if ok_rows.any() else 0.0
    stats["z_bitexact_frac"] = float((Z == Z_ref).mean())
    stats["z_max_over_tol"] = float(np.nanmax((dz / tol)[ok_rows])) if ok_rows.any() else 0.0
    if bad.any():
        fails.append("%d Z entries outside tolerance (worst ratio %.2f)" % (int(bad.sum()), stats["z_max_over_tol"]))
    fin = np.isfinite(lPz_ref) & ok_rows
    dl = np.abs(lPz - lPz_ref)[fin] / np.maximum(1.0, np.abs(lPz_ref[fin]))
    stats["lpz_max_rel"] = float(dl.max()) if dl.size else 0.0
    ltol = lpz_tolerance(lPz_ref, cond, lsens)[fin]
    lratio = np.abs(lPz - lPz_ref)[fin] / ltol
    stats["lpz_max_over_tol"] = float(lratio.max()) if lratio.size else 0.0
    if lratio.size and not (lratio <= 1.0).all():
        fails.append("%d lPz entries outside tolerance (max relative difference %.2e, worst ratio %.2f)" % (int((lratio > 1.0).sum()), dl.max(), lratio.max()))
    same_nonfinite = np.array_equal(np.isfinite(lPz_ref), np.isfinite(lPz) | ~ok_rows | ~np.isfinite(lPz_ref))
    if not same_nonfinite:
        fails.append("non-finite lPz pattern differs")
    return stats, fails
