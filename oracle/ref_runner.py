"""Runs the UNMODIFIED reference from oracle/_ref (see oracle/build_ref.py) on descriptor / coordinate arrays.

Test infrastructure only -- see oracle/__init__.py.  bench.py's CPU arms time `pair_geometry_ref`, which is exactly
what the reference's video loop does per pair with precomputed features (video_processing.py:73-80):

    KeyPoints(q).match_static_kps(KeyPoints(t))            evenvizion/processing/matching.py:131-163
    compute_homography(static_a, static_b)                 evenvizion/processing/utils.py:328-363
"""
import os
import sys
import types

_REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_mods = None


def available():
    return os.path.isdir(os.path.join(_REF_DIR, "evenvizion", "processing"))


def load():
    """Import the reference's processing modules from oracle/_ref.  The package __init__ files pull in the examples
    and visualisation layers; empty namespace packages are registered instead so that only
    evenvizion.processing.{constants, utils, matching} are executed -- those three files run unmodified."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise ImportError("oracle/_ref is missing: run `python oracle/build_ref.py` where /root/reference exists")
    import cv2
    if not hasattr(cv2, "xfeatures2d"):
        cv2.xfeatures2d = types.SimpleNamespace(SIFT_create=cv2.SIFT_create)
    if _REF_DIR not in sys.path:
        sys.path.insert(0, _REF_DIR)                      # imutils stand-in
    import importlib.util
    for name, sub in (("evenvizion", "evenvizion"), ("evenvizion.processing", "evenvizion/processing")):
        if name not in sys.modules:
            pkg = types.ModuleType(name)
            pkg.__path__ = [os.path.join(_REF_DIR, sub)]
            sys.modules[name] = pkg
    out = {}
    for m in ("constants", "utils", "matching"):
        full = "evenvizion.processing." + m
        if full not in sys.modules:
            spec = importlib.util.spec_from_file_location(full, os.path.join(_REF_DIR, "evenvizion", "processing", m + ".py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[full] = mod
            spec.loader.exec_module(mod)
        out[m] = sys.modules[full]
    _mods = out
    return out


def pair_geometry_ref(q_coords, q_desc, t_coords, t_desc):
    """The reference's own per-pair geometry; returns H (3x3) or None (NoMatchesException / HomographyException)."""
    import numpy as np
    m = load()
    KeyPoints, NoMatches = m["matching"].KeyPoints, m["matching"].NoMatchesException
    try:
        sa, sb = KeyPoints(q_coords, np.asarray(q_desc, np.float32)).match_static_kps(
            KeyPoints(t_coords, np.asarray(t_desc, np.float32)))
        return m["utils"].compute_homography(sa, sb)
    except (NoMatches, m["utils"].HomographyException):
        return None
    except Exception:        # cv2.error when fewer than 4 points reach findHomography (uncaught in the reference)
        return None


def _worker_init():
    import cv2
    cv2.setNumThreads(1)
    load()


def _worker(args):
    return pair_geometry_ref(*args) is not None


def time_pairs(frames, n_workers):
    """frames: list of (coords, desc); times the pairs (k+1, k) over a pool of n_workers single-threaded OpenCV worker
    processes, each running the unmodified reference.  Returns (seconds, n_pairs, n_ok)."""
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
