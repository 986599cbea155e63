"""Oracle: descriptor matching (reference `evenvizion/processing/matching.py`).

Test infrastructure only -- see oracle/__init__.py.
"""
import numpy as np

LOWES_RATIO = 0.5              # reference constants.py:25
MINIMUM_MATCHING_POINTS = 4    # reference constants.py:28


def knn_top2(q_desc, t_desc):
    """Exact 2-NN under L2, restating `cv2.BFMatcher(NORM_L2).knnMatch(q, t, 2)`
    as called at reference matching.py:102-108.

    SURVEY P1: for integer-valued descriptors <= 255 OpenCV's result equals a
    stable arg-sort of the exact integer squared distance (ties -> lowest train
    index, also for the second neighbour) and `distance == sqrtf((float)d2)`.

    Returns idx (Nq,2) int32 (-1 where absent), d2 (Nq,2) int64 (-1 where absent).
    """
    q = np.ascontiguousarray(q_desc).astype(np.float32)
    t = np.ascontiguousarray(t_desc).astype(np.float32)
    nq, nt = q.shape[0], t.shape[0]
    idx = np.full((nq, 2), -1, np.int32)
    d2o = np.full((nq, 2), -1, np.int64)
    if nq == 0 or nt == 0:
        return idx, d2o
    # all partial sums are non-negative integers < 2^24 -> f32 BLAS is exact
    qn = (q.astype(np.int64) ** 2).sum(1)
    tn = (t.astype(np.int64) ** 2).sum(1)
    step = max(1, (1 << 24) // max(nt, 1))
    for s in range(0, nq, step):
        e = min(nq, s + step)
        dot = (q[s:e] @ t.T).astype(np.int64)
        d2 = qn[s:e, None] + tn[None, :] - 2 * dot
        key = d2 * np.int64(1 << 20) + np.arange(nt, dtype=np.int64)[None, :]  # (d2, idx) lexicographic
        if nt >= 2:
            part = np.partition(key, 1, axis=1)[:, :2]
            part.sort(axis=1)
            idx[s:e] = (part & ((1 << 20) - 1)).astype(np.int32)
            d2o[s:e] = part >> 20
        else:
            idx[s:e, 0] = 0
            d2o[s:e, 0] = d2[:, 0]
    return idx, d2o


def ratio_survivors(idx, d2, ratio=LOWES_RATIO):
    """reference matching.py:190: `len(matches)==2 and m0.distance < m1.distance*ratio`
    where distance is an f32 (`sqrtf((float)d2)`) widened to a Python double.
    SURVEY P2: this is NOT `4*d2_0 < d2_1`."""
    have2 = idx[:, 1] >= 0
    d0 = np.sqrt(np.where(have2, d2[:, 0], 0).astype(np.float32)).astype(np.float64)
    d1 = np.sqrt(np.where(have2, d2[:, 1], 0).astype(np.float32)).astype(np.float64)
    return have2 & (d0 < d1 * float(ratio))


def collision_filter(idx, surv):
    """reference matching.py:166-239 (`lowes_ratio_test` + `filter_corresponding_points`):
    every train index claimed by more than one surviving query is dropped together
    with all of its claimants; output ordered by first insertion of the train index,
    which is ascending query index.  Returns list of (train, query)."""
    q = np.nonzero(surv)[0]
    t = idx[q, 0]
    if len(q) == 0:
        return np.zeros((0, 2), np.int32)
    cnt = np.bincount(t, minlength=int(t.max()) + 1)
    keep = cnt[t] == 1
    return np.stack([t[keep], q[keep]], 1).astype(np.int32)


def remove_double_matching(pts_a, pts_b):
    """reference utils.py:41-68: dict keyed by the exact (x, y) of pts_a; a repeated
    key keeps its FIRST position but the LAST value.  Returns (new_a, new_b, keep_pos, val_pos)
    where new_a = pts_a[keep_pos], new_b = pts_b[val_pos]."""
    pts_a = np.asarray(pts_a, np.float32).reshape(-1, 2)
    pts_b = np.asarray(pts_b, np.float32).reshape(-1, 2)
    first = {}
    last = {}
    for i in range(len(pts_a)):
        k = (float(pts_a[i, 0]), float(pts_a[i, 1]))   # -0.0 == 0.0 hash-equal, as in Python
        if k not in first:
            first[k] = i
        last[k] = i
    keep = np.fromiter(first.values(), np.int64, len(first))
    val = np.fromiter((last[k] for k in first), np.int64, len(first))
    return pts_a[keep], pts_b[val], keep, val


def match_kps(q_coords, q_desc, t_coords, t_desc, ratio=LOWES_RATIO,
              min_matching_pts=MINIMUM_MATCHING_POINTS):
    """reference matching.py:75-129 (`KeyPoints.match_kps`): self = query, acceding = train.
    Returns dict(status, pts_a, pts_b, matches, idx, d2, surv).  status: 0 ok,
    1 = NoMatchesException("len(matches) < min_matching_pts")."""
    idx, d2 = knn_top2(q_desc, t_desc)
    surv = ratio_survivors(idx, d2, ratio)
    m = collision_filter(idx, surv)
    out = dict(idx=idx, d2=d2, surv=surv, matches=m, status=0,
               pts_a=np.zeros((0, 2), np.float32), pts_b=np.zeros((0, 2), np.float32))
    if len(m) < min_matching_pts:
        out["status"] = 1
        return out
    pa = np.asarray(q_coords, np.float32)[m[:, 1]]
    pb = np.asarray(t_coords, np.float32)[m[:, 0]]
    pa, pb, keep, val = remove_double_matching(pa, pb)
    out.update(pts_a=pa, pts_b=pb, keep=keep, val=val)
    return out
