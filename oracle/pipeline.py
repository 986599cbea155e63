"""Oracle: the per-pair geometry pipeline and the video-level chain, restating
reference matching.py:131-163 (`match_static_kps`), frame_processing.py:73-108,
utils.py:328-363 (`compute_homography`) and video_processing.py:67-107.

Test infrastructure only -- see oracle/__init__.py.
"""
import numpy as np
from . import matching, ransac, static_filter, chain

LENGTH_ACCOUNTED_POINTS = 0.7   # reference constants.py:19

# per-pair status codes (mirrored by include/evz.h)
ST_OK = 0
ST_FEW_MATCHES = 1      # NoMatchesException("len(matches) < min_matching_pts")   matching.py:113-116
ST_FEW_POINTS = 3       # < 4 points reach findHomography (cv2.error in the reference)
ST_NO_MODEL_1 = 4       # RANSAC #1 found no model -> NoMatchesException           matching.py:158-159
ST_NO_MODEL_2 = 5       # RANSAC #2 found no model -> HomographyException          utils.py:361-362
ST_FEW_INLIERS = 6      # sum(mask) < 0.7 len(mask) -> HomographyException         utils.py:359-360


def pair_geometry(q_coords, q_desc, t_coords, t_desc, n_hyp=1024, seed=0, pair_id=0,
                  ratio=0.5, thresh=3.0, S_prev=None):
    """One frame pair: self/query = the NEW frame, acceding/train = the PREVIOUS frame.
    Returns dict with every intermediate the CUDA path exposes."""
    out = dict(status=ST_OK, H=None)
    mk = matching.match_kps(q_coords, q_desc, t_coords, t_desc, ratio)
    out["match"] = mk
    if mk["status"] != 0:
        out["status"] = ST_FEW_MATCHES
        return out
    pa, pb = mk["pts_a"], mk["pts_b"]
    r1 = ransac.find_homography_seeded(pa, pb, n_hyp, seed, pair_id, 1, thresh)
    out["ransac1"] = r1
    if r1["status"] != 0:
        out["status"] = ST_FEW_POINTS if r1["status"] == ransac.ST_TOO_FEW else ST_NO_MODEL_1
        return out
    keep, best_r, bad = static_filter.static_points(r1["H"], pa, pb)
    out["static_keep"] = keep
    out["static_r"] = best_r
    sa, sb = pa[keep], pb[keep]
    # frame_processing.py:102-104 de-duplicates again (a no-op for a single feature type)
    sa, sb, _, _ = matching.remove_double_matching(sa, sb)
    out["static_a"], out["static_b"] = sa, sb
    if S_prev is not None:                                  # utils.py:351-355
        S_prev = np.asarray(S_prev, np.float64)
        sa2 = np.array([chain.homography_transformation(p, S_prev) for p in sa]).reshape(-1, 2)
        sb2 = np.array([chain.homography_transformation(p, S_prev) for p in sb]).reshape(-1, 2)
    else:
        sa2, sb2 = sa, sb
    r2 = ransac.find_homography_seeded(sa2, sb2, n_hyp, seed, pair_id, 2, thresh)
    out["ransac2"] = r2
    if r2["status"] != 0:
        out["status"] = ST_FEW_POINTS if r2["status"] == ransac.ST_TOO_FEW else ST_NO_MODEL_2
        return out
    if int(r2["mask"].sum()) < LENGTH_ACCOUNTED_POINTS * len(r2["mask"]):
        out["status"] = ST_FEW_INLIERS
        return out
    out["H"] = r2["H"]
    return out


def video_chain(frames, n_hyp=1024, seed=0, none_h_processing=True, reference_exact=False,
                ratio=0.5, thresh=3.0):
    """frames: list of (coords (N,2) f32, desc (N,D) u8).  Pair p = (new frame p+1, old frame p).
    parallel mode (default): every pair independent in the frame plane, then forward-fill +
    prefix product.  reference_exact: RANSAC #2 runs on points pre-transformed by the running
    superposition, serially, exactly as video_processing.py:67-105 does.
    Returns dict(G or H list, valid, S (F,3,3), status list)."""
    P = len(frames) - 1
    status = np.zeros(P, np.int32)
    if not reference_exact:
        G = np.tile(np.eye(3), (P, 1, 1))
        valid = np.zeros(P, bool)
        for p in range(P):
            r = pair_geometry(frames[p + 1][0], frames[p + 1][1], frames[p][0], frames[p][1],
                              n_hyp, seed, p, ratio, thresh)
            status[p] = r["status"]
            if r["H"] is not None:
                G[p] = r["H"]
                valid[p] = True
        Gf = chain.fill_none(G, valid, none_h_processing)
        S = chain.chain_products(Gf)
        return dict(G=G, valid=valid, S=S, status=status, H_fixed=chain.fixed_plane_H(S))
    Hs = []
    valid = np.zeros(P, bool)
    S = None
    Hprev = None
    sup = [np.eye(3)]
    first = True
    for p in range(P):
        r = pair_geometry(frames[p + 1][0], frames[p + 1][1], frames[p][0], frames[p][1],
                          n_hyp, seed, p, ratio, thresh, S_prev=S)
        status[p] = r["status"]
        H = r["H"]
        valid[p] = H is not None
        if H is None:
            H = Hprev if none_h_processing else None
        Hs.append(H)
        S = chain.matrix_superposition(H, S, first) if H is not None or S is not None else S
        if H is not None:
            first = False
            Hprev = H
        sup.append(np.eye(3) if S is None else np.asarray(S, np.float64))
    return dict(H_fixed=Hs, valid=valid, S=np.array(sup), status=status)


def pair_static_multi(frames_q, frames_t, n_hyp=1024, seed=0, pair_id=0, ratio=0.5, thresh=3.0):
    """`FrameProcessing.concatenate_all_features_types` (frame_processing.py:91-104) for one pair:
    frames_q / frames_t are lists over feature types of (coords, desc) of the new / previous frame.
    Every type goes through match_kps -> RANSAC #1 -> static filter (same pair_id and level 1 for
    every type); the static points are concatenated in type order and de-duplicated.
    Returns (status, static_a, static_b, per-type intermediates)."""
    all_a, all_b, per = [], [], []
    for (qc, qd), (tc, td) in zip(frames_q, frames_t):
        if qd is None or td is None:
            return ST_FEW_MATCHES, None, None, per         # matching.py:104-107: NoMatchesException
        mk = matching.match_kps(qc, qd, tc, td, ratio)
        if mk["status"] != 0:
            return ST_FEW_MATCHES, None, None, per
        r1 = ransac.find_homography_seeded(mk["pts_a"], mk["pts_b"], n_hyp, seed, pair_id, 1, thresh)
        if r1["status"] != 0:
            return (ST_FEW_POINTS if r1["status"] == ransac.ST_TOO_FEW else ST_NO_MODEL_1), None, None, per
        keep, best_r, _ = static_filter.static_points(r1["H"], mk["pts_a"], mk["pts_b"])
        per.append(dict(match=mk, ransac1=r1, keep=keep))
        all_a.append(mk["pts_a"][keep]); all_b.append(mk["pts_b"][keep])
    sa, sb, _, _ = matching.remove_double_matching(np.concatenate(all_a), np.concatenate(all_b))
    return ST_OK, sa, sb, per


def video_chain_multi(feats, n_hyp=1024, seed=0, none_h_processing=True, reference_exact=False,
                      ratio=0.5, thresh=3.0):
    """`get_homography_dict` (video_processing.py:67-105) over several feature types.
    feats: {type: [(coords, desc) per frame]}.  Same return value as video_chain, plus the merged
    static sets per pair (`static`)."""
    types = list(feats)
    F = len(feats[types[0]])
    P = F - 1
    status = np.zeros(P, np.int32)
    static = []
    G = np.tile(np.eye(3), (P, 1, 1))
    valid = np.zeros(P, bool)
    Hs, sup = [], [np.eye(3)]
    S, Hprev, first = None, None, True
    for p in range(P):
        st, sa, sb, _ = pair_static_multi([feats[t][p + 1] for t in types], [feats[t][p] for t in types],
                                          n_hyp, seed, p, ratio, thresh)
        static.append((sa, sb))
        H = None
        if st == ST_OK:
            if reference_exact and S is not None:
                Sp = np.asarray(S, np.float64)
                sa2 = np.array([chain.homography_transformation(q, Sp) for q in sa]).reshape(-1, 2)
                sb2 = np.array([chain.homography_transformation(q, Sp) for q in sb]).reshape(-1, 2)
            else:
                sa2, sb2 = sa, sb
            r2 = ransac.find_homography_seeded(sa2, sb2, n_hyp, seed, p, 2, thresh)
            if r2["status"] != 0:
                st = ST_FEW_POINTS if r2["status"] == ransac.ST_TOO_FEW else ST_NO_MODEL_2
            elif int(r2["mask"].sum()) < LENGTH_ACCOUNTED_POINTS * len(r2["mask"]):
                st = ST_FEW_INLIERS
            else:
                H = r2["H"]
        status[p] = st
        valid[p] = H is not None
        if not reference_exact:
            if H is not None:
                G[p] = H
            continue
        if H is None:
            H = Hprev if none_h_processing else None
        Hs.append(H)
        S = chain.matrix_superposition(H, S, first) if H is not None or S is not None else S
        if H is not None:
            first = False
            Hprev = H
        sup.append(np.eye(3) if S is None else np.asarray(S, np.float64))
    if not reference_exact:
        Gf = chain.fill_none(G, valid, none_h_processing)
        Sx = chain.chain_products(Gf)
        return dict(G=G, valid=valid, S=Sx, status=status, H_fixed=chain.fixed_plane_H(Sx), static=static)
    return dict(H_fixed=Hs, valid=valid, S=np.array(sup), status=status, static=static)
