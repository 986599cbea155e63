"""CPU baseline: the reference's own geometry path, restated call for call with live OpenCV.

Test infrastructure only -- see oracle/__init__.py.  Used by bench.py's `cpu_baseline` leg and
by `bench.py --impl reference` (the reference itself is pure Python and lives outside the repo,
so it cannot travel to the GPU box; this module makes the same OpenCV calls in the same order
with the same per-point Python glue, so its cost profile is the reference's).

  pair_geometry_cv2(q, t)  ==  KeyPoints(q).match_static_kps(KeyPoints(t))      matching.py:131-163
                               + compute_homography(static_a, static_b)          utils.py:328-363
"""
import numpy as np

LOWES_RATIO = 0.5
MINIMUM_MATCHING_POINTS = 4
THRESHOLD_FOR_FIND_HOMOGRAPHY = 3.0
LENGTH_ACCOUNTED_POINTS = 0.7


def _match_kps(cv2, q_coords, q_desc, t_coords, t_desc):
    matcher = cv2.DescriptorMatcher_create("BruteForce")
    raw = matcher.knnMatch(q_desc, t_desc, 2)
    train, query = {}, {}
    for m in raw:                                           # matching.py:189-197
        if len(m) == 2 and m[0].distance < m[1].distance * LOWES_RATIO:
            train.setdefault(m[0].trainIdx, []).append(m[0].queryIdx)
            query.setdefault(m[0].queryIdx, []).append(m[0].trainIdx)
    dead = [k for k, v in train.items() if len(v) > 1]      # matching.py:226-236
    for k in set(dead):
        del train[k]
    matches = [(t, q[0]) for t, q in train.items()]
    if len(matches) < MINIMUM_MATCHING_POINTS:
        return None
    pts_a = np.float32([q_coords[i] for (_, i) in matches])
    pts_b = np.float32([t_coords[i] for (i, _) in matches])
    d = {}
    for i in range(len(pts_a)):                             # utils.py:63-67
        d[(pts_a[i][0], pts_a[i][1])] = pts_b[i]
    return [np.array(k) for k in d], list(d.values())


def _static(cv2, pts_a, pts_b):
    H, _ = cv2.findHomography(np.array(pts_a), np.array(pts_b), cv2.RANSAC, THRESHOLD_FOR_FIND_HOMOGRAPHY)
    if H is None:
        return None
    groups = {}
    for i in range(len(pts_a)):                             # utils.py:316-324
        v = np.dot(H, (pts_a[i][0], pts_a[i][1], 1))
        r = round(np.sum(np.subtract(v[:2] / v[2], pts_b[i]) ** 2) ** 0.5)
        groups.setdefault(r, []).append(i)
    best, n = None, 0
    for k, v in groups.items():
        if len(v) > n:
            n, best = len(v), k
    keep = groups[best]
    return np.array([pts_a[i] for i in keep]), np.array([pts_b[i] for i in keep])


def pair_geometry_cv2(q_coords, q_desc, t_coords, t_desc):
    """Returns the 3x3 homography (new frame -> previous frame) or None."""
    import cv2
    m = _match_kps(cv2, q_coords, np.asarray(q_desc, np.float32), t_coords, np.asarray(t_desc, np.float32))
    if m is None or len(m[0]) < 4:
        return None
    s = _static(cv2, *m)
    if s is None or len(s[0]) < 4:
        return None
    H, status = cv2.findHomography(np.array(s[0]), np.array(s[1]), cv2.RANSAC, THRESHOLD_FOR_FIND_HOMOGRAPHY)
    if H is None or np.sum(status) < LENGTH_ACCOUNTED_POINTS * len(status):
        return None
    return H


def _worker_init():
    import cv2
    cv2.setNumThreads(1)


def _worker(args):
    return pair_geometry_cv2(*args) is not None


def time_pairs(frames, n_workers):
    """frames: list of (coords, desc); times pairs (k+1, k) over a process pool of n_workers
    single-threaded OpenCV workers.  Returns (seconds, n_pairs, n_ok)."""
    import time
    import multiprocessing as mp
    jobs = [(frames[k + 1][0], frames[k + 1][1], frames[k][0], frames[k][1]) for k in range(len(frames) - 1)]
    if n_workers <= 1:
        _worker_init()
        t = time.perf_counter()
        ok = sum(_worker(j) for j in jobs)
        return time.perf_counter() - t, len(jobs), ok
    ctx = mp.get_context("spawn")      # the parent may hold a CUDA context
    with ctx.Pool(n_workers, initializer=_worker_init) as pool:
        pool.map(_worker, jobs[:n_workers])                 # warm the workers
        t = time.perf_counter()
        ok = sum(pool.map(_worker, jobs, chunksize=1))
        dt = time.perf_counter() - t
    return dt, len(jobs), ok
