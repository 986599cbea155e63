"""Oracle: None-H fallback, cumulative superposition, object-coordinate remap
(reference video_processing.py:94-105, utils.py:71-145,184-211, fixed_coordinate_system.py:19-122).

Test infrastructure only -- see oracle/__init__.py.
"""
import numpy as np


def matrix_superposition(H, S, first=False):
    """reference utils.py:118-145."""
    if H is not None:
        if first:
            S = H
        else:
            S = np.dot(H, S)
            S = np.divide(S, S[2][2])
    return S


def superposition_dict(homography_dict):
    """reference utils.py:184-211 (left fold, frame 1 -> identity)."""
    out = {1: [[1, 0, 0], [0, 1, 0], [0, 0, 1]]}
    S = None
    first = True
    for frame_no, frame_H in homography_dict.items():
        S = matrix_superposition(frame_H["H"], S, first)
        first = False
        out[frame_no] = S
    return out


def homography_transformation(vec, H):
    """reference utils.py:71-92."""
    v = np.append(np.asarray(vec, np.float64), [1.0])[:3]
    nv = np.dot(np.asarray(H, np.float64), v)
    return nv[:-1] / nv[-1]


def fill_none(G, valid, policy_prev=True):
    """None-H policy on frame-plane matrices G (P,3,3).  policy_prev=True: reuse the previous
    pair's matrix (reference video_processing.py:95-96; equivalent in the frame plane, see
    DESIGN.md); False: identity step.  Leading invalid pairs -> identity (the reference crashes)."""
    G = np.array(G, np.float64).reshape(-1, 3, 3)
    out = np.empty_like(G)
    prev = np.eye(3)
    for k in range(len(G)):
        if valid[k]:
            prev = G[k]
            out[k] = G[k]
        else:
            out[k] = prev if policy_prev else np.eye(3)
    return out


def chain_products(G):
    """S_k = normalise(S_{k-1} . G_k), S_0 = I: frame-plane step matrices (new frame -> previous
    frame) to cumulative superposition (frame k -> frame 1).  Returns (P+1,3,3) with S[0]=I."""
    G = np.asarray(G, np.float64).reshape(-1, 3, 3)
    S = np.empty((len(G) + 1, 3, 3))
    S[0] = np.eye(3)
    for k in range(len(G)):
        T = S[k] @ G[k]
        S[k + 1] = T / T[2, 2]
    return S


def fixed_plane_H(S):
    """H_k = S_k . S_{k-1}^{-1} (normalised): the matrices the reference stores in
    dict_with_homography_matrix.json, so that its `superposition_dict` (a LEFT fold)
    reproduces S."""
    out = np.empty((len(S) - 1, 3, 3))
    for k in range(1, len(S)):
        T = S[k] @ np.linalg.inv(S[k - 1])
        out[k - 1] = T / T[2, 2]
    return out


def from_original_to_fix(original_coordinates, sup, original_shape, resize_shape):
    """reference fixed_coordinate_system.py:19-69."""
    oh, ow = original_shape
    rh, rw = resize_shape
    hc = int(rh) / oh
    wc = int(rw) / ow
    out = {}
    for frame_no, rects in original_coordinates.items():
        out[frame_no] = []
        for rect in rects:
            nr = dict(rect)
            nr["x1"], nr["y1"] = np.around(
                homography_transformation([wc * rect["x1"], hc * rect["y1"]], sup[frame_no]), decimals=2)
            out[frame_no].append(nr)
    return out


def from_fix_to_original(fix_coordinates, sup, original_shape, resize_shape):
    """reference fixed_coordinate_system.py:72-122 (scales BEFORE the inverse transform, as the
    reference does)."""
    oh, ow = original_shape
    rh, rw = resize_shape
    hc = oh / rh
    wc = ow / rw
    out = {}
    for frame_no, rects in fix_coordinates.items():
        out[frame_no] = []
        Hi = np.linalg.inv(np.asarray(sup[frame_no], np.float64))
        for rect in rects:
            nr = dict(rect)
            nr["x1"], nr["y1"] = np.around(
                homography_transformation([wc * rect["x1"], hc * rect["y1"]], Hi), decimals=2)
            out[frame_no].append(nr)
    return out


def max_movement(sup, h, w):
    """reference processing_visualization.py:404-418 (`heatmap_video_processing`): the maximum
    transformed coordinate over the h x w pixel grid and all frames except the last dict entry."""
    ys, xs = np.mgrid[0:h, 0:w]
    p = np.stack([xs.ravel(), ys.ravel(), np.ones(h * w)], 0).astype(np.float64)
    best = -np.inf
    keys = list(sup.keys())
    for k in keys[:-1]:
        q = np.asarray(sup[k], np.float64) @ p
        best = max(best, float((q[:2] / q[2]).max()))
    return best
