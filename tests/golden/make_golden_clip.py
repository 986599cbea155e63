#!/usr/bin/env python
"""Golden vectors of the WHOLE bundled clip (BASELINE config 1): SIFT and ORB features of all 121
frames, and what the UNMODIFIED reference (/root/reference, imported read-only) produces from them.

    python tests/golden/make_golden_clip.py      ->  tests/golden/clip_full.npz

Contents:
  sift{f}_c / sift{f}_d, orb{f}_c / orb{f}_d   features of frame f (cv2 4.13 in the build container; the
                                               frames are resized to width 400 like the reference does)
  ref_H[120,3,3], ref_resize_info              the reference's own get_homography_dict(capture, 400) with
                                               features ["SIFT", "ORB"] (its default list minus non-free SURF)
  orbmk{p}_pts_a / _pts_b                      the reference's KeyPoints.match_kps on the ORB features of
                                               pairs 0..11 (bit-exact pin of the D = 32 matcher path)
  ref_cat{p}_a / _b                            the reference's concatenate_all_features_types output for
                                               pairs 0..11 (cv2's own RANSAC sampling: agreement, not identity)
  frame{f}                                     frames 0..5 resized to 400 x 224 (u8 BGR) for the
                                               get_homography_dict test on a capture object
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import install_shims, REF     # noqa: E402


def main():
    install_shims()
    import imutils
    from evenvizion.processing import matching as rm
    from evenvizion.processing import frame_processing as rfp
    from evenvizion.processing import video_processing as rvp

    types = ["SIFT", "ORB"]
    # the reference's default list contains the non-free SURF; get_homography_dict never passes a list
    # (video_processing.py:69,74), so the default is narrowed for this run
    orig_init = rfp.FrameProcessing.__init__

    def init(self, frame, features_type_list=None):
        orig_init(self, frame, features_type_list or list(types))
    rfp.FrameProcessing.__init__ = init

    video = os.path.join(REF, "evenvizion/examples/test_video/test_video.mp4")
    out = {}
    cap = cv2.VideoCapture(video)
    frames = []
    ok, img = cap.read()
    while ok:
        frames.append(imutils.resize(img, width=400))
        ok, img = cap.read()
    assert len(frames) == 121, len(frames)
    feats = {t: [] for t in types}
    for f, fr in enumerate(frames):
        fp = rfp.FrameProcessing(fr)
        for t in types:
            c, d = fp.detect_and_describe_features(t)
            if t == "SIFT":
                assert (d == np.rint(d)).all() and d.max() <= 255 and d.min() >= 0
            d = d.astype(np.uint8)
            feats[t].append((c, d))
            out[f"{t.lower()}{f}_c"] = c
            out[f"{t.lower()}{f}_d"] = d
        if f < 6:
            out[f"frame{f}"] = fr
    out["n_frames"] = np.int32(len(frames))

    hd = rvp.get_homography_dict(cv2.VideoCapture(video), 400, None, True)
    out["ref_H"] = np.array([hd[k]["H"] for k in range(2, 122)], np.float64)
    out["ref_resize_info"] = np.array([hd["resize_info"]["h"], hd["resize_info"]["w"]], np.int32)

    for p in range(12):
        kq = rm.KeyPoints(feats["ORB"][p + 1][0], feats["ORB"][p + 1][1])
        kt = rm.KeyPoints(feats["ORB"][p][0], feats["ORB"][p][1])
        pa, pb = kq.match_kps(kt)
        out[f"orbmk{p}_pts_a"] = np.array(pa, np.float32).reshape(-1, 2)
        out[f"orbmk{p}_pts_b"] = np.array(pb, np.float32).reshape(-1, 2)
        a, b = rfp.FrameProcessing(frames[p + 1]).concatenate_all_features_types(rfp.FrameProcessing(frames[p]))
        out[f"ref_cat{p}_a"] = np.array(a, np.float32).reshape(-1, 2)
        out[f"ref_cat{p}_b"] = np.array(b, np.float32).reshape(-1, 2)
    path = os.path.join(HERE, "clip_full.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
