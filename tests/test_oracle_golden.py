"""The oracle against the golden vectors produced by the reference itself
(tests/golden/make_golden.py).  CPU only."""
import numpy as np

from oracle import chain, matching, ransac, static_filter


def test_knn_top2_matches_opencv_bruteforce(golden):
    for i in range(int(golden["knn_n"])):
        q, t = golden[f"knn{i}_q"], golden[f"knn{i}_t"]
        idx, d2 = matching.knn_top2(q, t)
        assert (idx == golden[f"knn{i}_idx"]).all(), i
        dist = np.where(idx >= 0, np.sqrt(np.maximum(d2, 0).astype(np.float32)), np.float32(-1))
        assert (dist == golden[f"knn{i}_dist"]).all(), i          # bit-exact f32 sqrt (P1)
        m = matching.collision_filter(idx, matching.ratio_survivors(idx, d2))
        assert (m == golden[f"knn{i}_lowe"]).all(), i              # ratio test + collision filter (P2, P6)


def test_ratio_edge_is_not_integer_compare(golden):
    # knn4 is the s1 = 4*s0 + 1 edge at s0 = 1049300: the reference REJECTS it, 4*s0 < s1 accepts
    idx, d2 = matching.knn_top2(golden["knn4_q"], golden["knn4_t"])
    assert 4 * d2[0, 0] < d2[0, 1]
    assert not matching.ratio_survivors(idx, d2)[0]
    assert len(golden["knn4_lowe"]) == 0


def test_match_kps_matches_reference(golden):
    for i in range(int(golden["mk_n"])):
        r = matching.match_kps(golden[f"mk{i}_qc"], golden[f"mk{i}_qd"], golden[f"mk{i}_tc"], golden[f"mk{i}_td"])
        assert r["status"] == 0
        assert np.array_equal(r["pts_a"], golden[f"mk{i}_pts_a"])
        assert np.array_equal(r["pts_b"], golden[f"mk{i}_pts_b"])


def test_match_kps_on_bundled_clip(golden):
    n = int(golden["clip_n"])
    for p in range(n - 1):
        r = matching.match_kps(golden[f"clip{p+1}_c"], golden[f"clip{p+1}_d"], golden[f"clip{p}_c"], golden[f"clip{p}_d"])
        assert np.array_equal(r["pts_a"], golden[f"clip{p}_pts_a"]), p
        assert np.array_equal(r["pts_b"], golden[f"clip{p}_pts_b"]), p


def test_remove_double_matching(golden):
    na, nb, _, _ = matching.remove_double_matching(golden["dd_a"], golden["dd_b"])
    assert np.array_equal(na, golden["dd_na"]) and np.array_equal(nb, golden["dd_nb"])


def _px(H, a):
    p = np.c_[a.astype(np.float64), np.ones(len(a))] @ np.asarray(H, np.float64).reshape(3, 3).T
    return p[:, :2] / p[:, 2:]


def test_refit_matches_findhomography_method0(golden):
    for i in range(int(golden["fh_n"])):
        a, b = golden[f"fh{i}_a"], golden[f"fh{i}_b"]
        H = ransac.refit(a, b)
        assert np.abs(_px(H, a) - _px(golden[f"fh{i}_H0"], a)).mean() < 1e-6      # <= 3e-8 px


def test_score_formula_bit_exact_against_opencv_mask(golden):
    for i in range(int(golden["fh_n"])):
        a, b2, Hr, mr = golden[f"fh{i}_a"], golden[f"fh{i}_b2"], golden[f"fh{i}_Hr"], golden[f"fh{i}_mr"]
        mask = ransac.reproj_err32(Hr.ravel(), a, b2) <= np.float32(9.0)
        assert np.array_equal(mask.astype(np.uint8), mr), i


def test_static_filter_matches_reference(golden):
    for i in range(int(golden["fh_n"])):
        a, b2, Hr = golden[f"fh{i}_a"], golden[f"fh{i}_b2"], golden[f"fh{i}_Hr"]
        keep, _, bad = static_filter.static_points(Hr, a, b2)
        assert not bad
        assert np.array_equal(a[keep], golden[f"fh{i}_static_a"])
        assert np.array_equal(b2[keep], golden[f"fh{i}_static_b"])


def test_seeded_ransac_agrees_with_opencv_on_inliers(golden):
    # reported agreement, not a bit-exact pin: OpenCV draws different samples
    for i in range(1, int(golden["fh_n"])):
        a, b2, mr = golden[f"fh{i}_a"], golden[f"fh{i}_b2"], golden[f"fh{i}_mr"]
        r = ransac.find_homography_seeded(a, b2, 1024, 0, i, 1)
        assert r["status"] == 0
        assert (r["mask"] == mr).mean() > 0.97
        inl = mr > 0
        assert np.abs(_px(r["H"], a[inl]) - _px(golden[f"fh{i}_Hr"], a[inl])).mean() < 0.25


def test_hypothesis_sampler_distinct_and_in_range():
    for m in (4, 5, 7, 100, 5000):
        idx = ransac.hyp_indices(7, 3, 2, 2048, m)
        s = np.sort(idx, 1)
        assert (s[:, 1:] != s[:, :-1]).all() and idx.min() >= 0 and idx.max() < m


def test_superposition_and_remap_known_answers(bundled):
    hd = {int(k): v for k, v in bundled["homography_dict"].items()}
    sup = chain.superposition_dict(hd)
    for k, v in bundled["superposition"].items():
        assert np.array_equal(np.asarray(sup[int(k)], np.float64), np.asarray(v, np.float64))
    ri = bundled["resize_info"]
    assert chain.max_movement(sup, ri["h"], ri["w"]) == bundled["max_movement"]     # metrics_file.txt
    oc = {int(k): v for k, v in bundled["original_coordinates"].items()}
    fixed = chain.from_original_to_fix(oc, sup, bundled["original_shape"], [ri["h"], ri["w"]])
    for k, rects in bundled["fixed_coordinates"].items():
        for r0, r1 in zip(rects, fixed[int(k)]):
            assert r0["x1"] == r1["x1"] and r0["y1"] == r1["y1"]
    back = chain.from_fix_to_original(fixed, sup, bundled["original_shape"], [ri["h"], ri["w"]])
    for k, rects in bundled["back_to_original"].items():
        for r0, r1 in zip(rects, back[int(k)]):
            assert abs(r0["x1"] - r1["x1"]) <= 0.011 and abs(r0["y1"] - r1["y1"]) <= 0.011


def test_parallel_chain_equals_reference_left_fold(bundled):
    # frame-plane G_k scanned on the right == fixed-plane H_k folded on the left
    hd = {int(k): v for k, v in bundled["homography_dict"].items()}
    sup = chain.superposition_dict(hd)
    keys = sorted(hd)
    S_ref = np.array([np.asarray(sup[k], np.float64) for k in [1] + keys])
    G = np.array([np.linalg.inv(S_ref[i]) @ S_ref[i + 1] for i in range(len(keys))])
    G /= G[:, 2:3, 2:3]
    S = chain.chain_products(G)
    assert np.abs(S - S_ref).max() < 1e-8
    Hf = chain.fixed_plane_H(S)
    Href = np.array([hd[k]["H"] for k in keys])
    assert np.abs(Hf - Href).max() < 1e-8


# ---------------------------------------------------------------- whole bundled clip, SIFT + ORB (BASELINE config 1)
def _clip():
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "clip_full.npz"))


def test_orb_match_kps_matches_reference():
    clip = _clip()
    for p in range(12):
        r = matching.match_kps(clip[f"orb{p+1}_c"], clip[f"orb{p+1}_d"], clip[f"orb{p}_c"], clip[f"orb{p}_d"])
        assert np.array_equal(r["pts_a"], clip[f"orbmk{p}_pts_a"]) and np.array_equal(r["pts_b"], clip[f"orbmk{p}_pts_b"]), p


def test_multi_type_chain_agrees_with_reference_get_homography_dict():
    """oracle.pipeline.video_chain_multi (seeded RANSAC) against the reference's own get_homography_dict on the bundled
    clip with ["SIFT", "ORB"] (cv2's own sampling): the merged static sets overlap almost entirely and the fixed-plane
    chains stay within a pixel of each other -- agreement, not identity (the hypotheses differ)."""
    from oracle import pipeline
    clip = _clip()
    n = 13
    feats = {t: [(clip[f"{t.lower()}{f}_c"], clip[f"{t.lower()}{f}_d"]) for f in range(n)] for t in ("SIFT", "ORB")}
    out = pipeline.video_chain_multi(feats, n_hyp=1024, seed=0, reference_exact=True)
    assert (out["status"] == 0).all()
    jac = []
    for p in range(n - 1):
        a = {tuple(x) for x in out["static"][p][0].tolist()}
        b = {tuple(x) for x in clip[f"ref_cat{p}_a"].tolist()}
        jac.append(len(a & b) / len(a | b))
    # (the static filter keeps one bin of ROUNDED displacements: a slightly different H1 moves points across the
    # rounding boundary, so the sets of two correct RANSAC runs overlap largely, not entirely)
    assert min(jac) > 0.5 and np.mean(jac) > 0.75, jac
    grid = np.stack(np.meshgrid(np.linspace(0, 400, 9), np.linspace(0, 224, 6)), -1).reshape(-1, 2)
    pg = np.c_[grid, np.ones(len(grid))]
    sup_ref = chain.superposition_dict({k + 2: {"H": clip["ref_H"][k].tolist()} for k in range(n - 1)})
    for k in range(n - 1):
        u = pg @ out["S"][k + 1].T; v = pg @ np.asarray(sup_ref[k + 2], np.float64).T
        assert np.abs(u[:, :2] / u[:, 2:] - v[:, :2] / v[:, 2:]).mean() < 1.0, k


def test_cpu_reference_port_reproduces_reference_H(golden):
    """bench.py's CPU arm (oracle/cpu_reference.py, kind "port") pinned to what the unmodified reference produced with
    the same deterministic OpenCV: compute_homography(match_static_kps(...)) on the synthetic pairs and the clip."""
    cv2 = __import__("pytest").importorskip("cv2")
    from oracle import cpu_reference
    for i in range(int(golden["mk_n"])):
        cv2.setRNGSeed(0)
        H = cpu_reference.pair_geometry_cv2(golden[f"mk{i}_qc"], golden[f"mk{i}_qd"], golden[f"mk{i}_tc"], golden[f"mk{i}_td"])
        assert H is not None and np.array_equal(H, golden[f"mk{i}_H"]), i
    for p in range(int(golden["clip_n"]) - 1):
        H = cpu_reference.pair_geometry_cv2(golden[f"clip{p+1}_c"], golden[f"clip{p+1}_d"], golden[f"clip{p}_c"], golden[f"clip{p}_d"])
        assert H is not None and np.array_equal(H, golden[f"clip{p}_H"]), p
