#!/usr/bin/env python
"""Generate the committed golden vectors by running the UNMODIFIED reference
(/root/reference, imported read-only) and OpenCV in the build container.

    python tests/golden/make_golden.py

The reference has no tests of its own, so these vectors are the
pin for the oracle: every array below is an output of reference code or of the OpenCV call
the reference makes, on inputs stored beside it.  /root/reference does not exist on the GPU
box, which is why the outputs are committed (small .npz / .json files).

Shims needed to import the reference here: a 10-line `imutils`
stand-in, `cv2.xfeatures2d.SIFT_create -> cv2.SIFT_create`; SURF is non-free and absent, so
feature lists are ["SIFT"].
"""
import json
import os
import sys
import types

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def install_shims():
    im = types.ModuleType("imutils")

    def resize(image, width=None, height=None, inter=cv2.INTER_AREA):
        h, w = image.shape[:2]
        if width is None and height is None:
            return image
        if width is None:
            r = height / float(h)
            dim = (int(w * r), height)
        else:
            r = width / float(w)
            dim = (width, int(h * r))
        return cv2.resize(image, dim, interpolation=inter)

    im.resize = resize
    im.is_cv3 = lambda or_better=False: True
    sys.modules["imutils"] = im
    if not hasattr(cv2, "xfeatures2d"):
        cv2.xfeatures2d = types.SimpleNamespace(SIFT_create=cv2.SIFT_create)
    sys.path.insert(0, REF)


def synth_pair(rng, n, d=128, corr=0.6, outl=0.2):
    g = rng.gamma(0.6, 1.0, (n, d))
    g /= np.linalg.norm(g, axis=1, keepdims=True)
    g = np.minimum(g, 0.2)
    g /= np.linalg.norm(g, axis=1, keepdims=True)
    da = np.clip(np.rint(512 * g), 0, 255).astype(np.uint8)
    ca = (rng.random((n, 2)) * [400, 224]).astype(np.float32)
    nc = int(corr * n)
    par = rng.permutation(n)[:nc]
    g2 = rng.gamma(0.6, 1.0, (n, d))
    g2 /= np.linalg.norm(g2, axis=1, keepdims=True)
    g2 = np.minimum(g2, 0.2)
    g2 /= np.linalg.norm(g2, axis=1, keepdims=True)
    db = np.clip(np.rint(512 * g2), 0, 255).astype(np.uint8)
    db[:nc] = np.clip(da[par].astype(np.int32) + np.rint(rng.normal(0, 6, (nc, d))), 0, 255).astype(np.uint8)
    th = rng.normal() * 0.01
    H = np.array([[np.cos(th), -np.sin(th), rng.normal() * 4], [np.sin(th), np.cos(th), rng.normal() * 4],
                  [rng.normal() * 1e-5, rng.normal() * 1e-5, 1]])
    p = np.c_[ca[par], np.ones(nc)] @ H.T
    cb = (rng.random((n, 2)) * [400, 224]).astype(np.float32)
    cc = (p[:, :2] / p[:, 2:] + rng.normal(0, 0.5, (nc, 2))).astype(np.float32)
    o = rng.random(nc) < outl
    cc[o] = cb[:nc][o]
    cb[:nc] = cc
    perm = rng.permutation(n)
    return ca, da, cb[perm], db[perm]


def main():
    install_shims()
    from evenvizion.processing import matching as rm
    from evenvizion.processing import utils as ru
    from evenvizion.processing import fixed_coordinate_system as rf
    from evenvizion.processing.frame_processing import FrameProcessing
    import imutils

    rng = np.random.default_rng(20261018)
    out = {}

    # ---- G1: knnMatch raw output on adversarial + synthetic descriptor sets (P1, P2)
    sets = []
    ca, da, cb, db = synth_pair(rng, 300)
    sets.append((db, da))
    q = rng.integers(0, 256, (64, 128)).astype(np.uint8)
    t = np.repeat(rng.integers(0, 256, (20, 128)).astype(np.uint8), 3, axis=0)   # exact ties in train
    sets.append((q, t))
    sets.append((np.zeros((5, 128), np.uint8), np.zeros((7, 128), np.uint8)))     # all-zero descriptors
    sets.append((rng.integers(0, 256, (9, 128)).astype(np.uint8), rng.integers(0, 256, (1, 128)).astype(np.uint8)))
    # ratio edge s1 = 4*s0 + 1 at large s0 (P2): q = 0, t0 has d2 = s0, t1 has d2 = 4 s0 + 1
    qe = np.zeros((1, 128), np.uint8)
    te = np.zeros((2, 128), np.uint8)
    te[0, :16] = 255; te[0, 16] = 94; te[0, 17] = 8      # 16*65025 + 8836 + 64 = 1049300
    s0 = int((te[0].astype(np.int64) ** 2).sum())
    assert s0 == 1049300
    rem = 4 * s0 + 1
    for j in range(128):
        v = min(255, int(np.sqrt(rem)))
        te[1, j] = v
        rem -= v * v
    assert rem == 0, rem
    sets.append((qe, te))
    sets.append((rng.integers(0, 256, (40, 32)).astype(np.uint8), rng.integers(0, 256, (50, 32)).astype(np.uint8)))  # ORB-width
    bf = cv2.DescriptorMatcher_create("BruteForce")
    for i, (qd, td) in enumerate(sets):
        raw = bf.knnMatch(qd.astype(np.float32), td.astype(np.float32), 2)
        idx = np.full((len(qd), 2), -1, np.int32)
        dist = np.full((len(qd), 2), -1, np.float32)
        for r, ms in enumerate(raw):
            for c, m in enumerate(ms):
                idx[r, c] = m.trainIdx
                dist[r, c] = m.distance
        lw = rm.lowes_ratio_test(raw)
        out[f"knn{i}_q"] = qd
        out[f"knn{i}_t"] = td
        out[f"knn{i}_idx"] = idx
        out[f"knn{i}_dist"] = dist
        out[f"knn{i}_lowe"] = np.array(lw, np.int32).reshape(-1, 2)
    out["knn_n"] = np.int32(len(sets))

    # ---- G2: KeyPoints.match_kps / match_static_kps on synthetic pairs (reference end to end)
    n_pairs = 3
    for i in range(n_pairs):
        ca, da, cb, db = synth_pair(rng, 350 + 50 * i)
        if i == 1:                      # duplicated query coordinates (multi-orientation SIFT keypoints)
            cb[10:20] = cb[0:10]
        kq = rm.KeyPoints(cb, db.astype(np.float32))    # self = new frame
        kt = rm.KeyPoints(ca, da.astype(np.float32))    # acceding = previous frame
        pa, pb = kq.match_kps(kt)
        sa, sb = kq.match_static_kps(kt)
        out[f"mk{i}_qc"] = cb; out[f"mk{i}_qd"] = db; out[f"mk{i}_tc"] = ca; out[f"mk{i}_td"] = da
        out[f"mk{i}_pts_a"] = np.array(pa, np.float32).reshape(-1, 2)
        out[f"mk{i}_pts_b"] = np.array(pb, np.float32).reshape(-1, 2)
        out[f"mk{i}_static_a"] = np.array(sa, np.float32).reshape(-1, 2)
        out[f"mk{i}_static_b"] = np.array(sb, np.float32).reshape(-1, 2)
        H = ru.compute_homography(sa, sb)
        out[f"mk{i}_H"] = H
    out["mk_n"] = np.int32(n_pairs)

    # ---- G3: findHomography pins: method 0 (DLT+LM) and RANSAC mask vs final H
    for i in range(4):
        n = [5, 40, 300, 1200][i]
        a = (rng.random((n, 2)) * [1920, 1080]).astype(np.float32)
        th = rng.normal() * 0.01
        Ht = np.array([[np.cos(th), -np.sin(th), rng.normal() * 8], [np.sin(th), np.cos(th), rng.normal() * 8],
                       [rng.normal() * 1e-5, rng.normal() * 1e-5, 1]])
        p = np.c_[a, np.ones(n)] @ Ht.T
        b = (p[:, :2] / p[:, 2:] + rng.normal(0, 0.5, (n, 2))).astype(np.float32)
        H0, _ = cv2.findHomography(a, b, 0)
        o = rng.random(n) < 0.35
        b2 = b.copy()
        b2[o] = (rng.random((int(o.sum()), 2)) * [1920, 1080]).astype(np.float32)
        Hr, mr = cv2.findHomography(a, b2, cv2.RANSAC, 3.0)
        out[f"fh{i}_a"] = a; out[f"fh{i}_b"] = b; out[f"fh{i}_H0"] = H0
        out[f"fh{i}_b2"] = b2; out[f"fh{i}_Hr"] = Hr; out[f"fh{i}_mr"] = mr.ravel()
        # displacement grouping with the reference's own functions
        d = ru.find_point_displacement(Hr, a, b2)
        ga, gb = ru.get_largest_group_points(d, a, b2)
        out[f"fh{i}_static_a"] = np.array(ga, np.float32).reshape(-1, 2)
        out[f"fh{i}_static_b"] = np.array(gb, np.float32).reshape(-1, 2)
    out["fh_n"] = np.int32(4)

    # ---- G4: remove_double_matching on a list with repeated keys
    pa = (rng.integers(0, 6, (60, 2))).astype(np.float32)
    pb = rng.random((60, 2)).astype(np.float32)
    na, nb = ru.remove_double_matching(pa, pb)
    out["dd_a"] = pa; out["dd_b"] = pb
    out["dd_na"] = np.array(na, np.float32).reshape(-1, 2); out["dd_nb"] = np.array(nb, np.float32).reshape(-1, 2)

    # ---- G5: bundled clip, first 13 frames: SIFT features (OpenCV CPU, out of scope) + reference path
    cap = cv2.VideoCapture(os.path.join(REF, "evenvizion/examples/test_video/test_video.mp4"))
    feats = []
    for f in range(13):
        ok, img = cap.read()
        assert ok
        img = imutils.resize(img, width=400)
        fp = FrameProcessing(img, ["SIFT"])
        c, d = fp.detect_and_describe_features("SIFT")
        assert (d == np.rint(d)).all() and d.max() <= 255
        feats.append((c, d.astype(np.uint8)))
        out[f"clip{f}_c"] = c
        out[f"clip{f}_d"] = d.astype(np.uint8)
    out["clip_n"] = np.int32(len(feats))
    out["clip_shape"] = np.array(img.shape[:2], np.int32)
    for p in range(len(feats) - 1):
        kq = rm.KeyPoints(feats[p + 1][0], feats[p + 1][1].astype(np.float32))
        kt = rm.KeyPoints(feats[p][0], feats[p][1].astype(np.float32))
        pa, pb = kq.match_kps(kt)
        sa, sb = kq.match_static_kps(kt)
        out[f"clip{p}_pts_a"] = np.array(pa, np.float32).reshape(-1, 2)
        out[f"clip{p}_pts_b"] = np.array(pb, np.float32).reshape(-1, 2)
        out[f"clip{p}_static_a"] = np.array(sa, np.float32).reshape(-1, 2)
        out[f"clip{p}_static_b"] = np.array(sb, np.float32).reshape(-1, 2)
        out[f"clip{p}_H"] = ru.compute_homography(sa, sb)       # frame-plane H by cv2's own RANSAC

    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)

    # ---- G6: scan + remap on the bundled JSONs (exact known-answer artefacts)
    hd, ri = ru.read_homography_dict(os.path.join(
        REF, "evenvizion/examples/test_video_processing/test_video/dict_with_homography_matrix.json"))
    sup = ru.superposition_dict(hd)
    oc = ru.read_json_with_coordinates(os.path.join(REF, "evenvizion/examples/test_video/original_coordinates.json"))
    fixed = rf.from_original_to_fix(oc, sup, [658, 1170], [ri["h"], ri["w"]])
    back = rf.from_fix_to_original(fixed, sup, [658, 1170], [ri["h"], ri["w"]])
    with open(os.path.join(HERE, "bundled_chain.json"), "w") as f:
        json.dump({
            "homography_dict": {str(k): v for k, v in hd.items()},
            "resize_info": ri,
            "original_shape": [658, 1170],
            "original_coordinates": {str(k): v for k, v in oc.items()},
            "superposition": {str(k): np.asarray(v, np.float64).tolist() for k, v in sup.items()},
            "fixed_coordinates": {str(k): [{kk: float(vv) for kk, vv in r.items()} for r in v] for k, v in fixed.items()},
            "back_to_original": {str(k): [{kk: float(vv) for kk, vv in r.items()} for r in v] for k, v in back.items()},
            "max_movement": 863.0428982580879,   # evenvizion/examples/test_video_processing/test_video/metrics_file.txt
        }, f)
    print("wrote", os.path.join(HERE, "reference_vectors.npz"), os.path.join(HERE, "bundled_chain.json"))


if __name__ == "__main__":
    main()
