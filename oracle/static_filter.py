"""Oracle: displacement-mode "static point" filter
(reference utils.py:289-325 `find_point_displacement`, utils.py:258-286 `get_largest_group_points`).

Test infrastructure only -- see oracle/__init__.py.
"""
import numpy as np

R_MAX = 16383   # displacement bins the CUDA kernel resolves exactly; larger / non-finite -> flagged


def displacement_bins(H, pts_a, pts_b):
    """r_i = round(||H*a_i - b_i||) with Python `round` (half-to-even) on f64
    (reference utils.py:317-320).  Operation order: (h0*x + h1*y) + h2, f64, no FMA.
    Returns (r int64 (M,), bad bool (M,)) where bad marks non-finite or > R_MAX values
    (the reference's `round()` raises on inf/nan)."""
    H = np.asarray(H, np.float64).reshape(3, 3)
    a = np.asarray(pts_a, np.float32).reshape(-1, 2).astype(np.float64)
    b = np.asarray(pts_b, np.float32).reshape(-1, 2).astype(np.float64)
    with np.errstate(all="ignore"):
        X = (H[0, 0] * a[:, 0] + H[0, 1] * a[:, 1]) + H[0, 2]
        Y = (H[1, 0] * a[:, 0] + H[1, 1] * a[:, 1]) + H[1, 2]
        W = (H[2, 0] * a[:, 0] + H[2, 1] * a[:, 1]) + H[2, 2]
        dx = X / W - b[:, 0]
        dy = Y / W - b[:, 1]
        dist = np.sqrt(dx * dx + dy * dy)
        bad = ~np.isfinite(dist) | (dist > R_MAX)
        r = np.rint(np.where(bad, 0.0, dist)).astype(np.int64)
    r[bad] = R_MAX + 1
    return r, bad


def largest_group(r):
    """Indices of the largest displacement group; ties -> the group whose key was inserted
    first into the dict (reference utils.py:279-282 uses a strict `>`), order preserved."""
    if len(r) == 0:
        return np.zeros(0, np.int64), -1
    vals, first, cnt = np.unique(r, return_index=True, return_counts=True)
    order = np.lexsort((first, -cnt))
    best = vals[order[0]]
    return np.nonzero(r == best)[0], int(best)


def static_points(H, pts_a, pts_b):
    r, bad = displacement_bins(H, pts_a, pts_b)
    keep, best = largest_group(r)
    return keep, best, bool(bad.any())
