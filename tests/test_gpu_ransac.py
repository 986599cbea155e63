"""K3 / K4 / K5 parity: seeded RANSAC, refit, static filter vs the oracle.  Needs a B200."""
import numpy as np
import pytest
import torch

from oracle import ransac, static_filter

pytestmark = pytest.mark.gpu


def _mk(rng, n, out_frac, noise=0.5):
    a = (rng.random((n, 2)) * [1920, 1080]).astype(np.float32)
    th = rng.normal() * 0.01
    s = 1 + rng.normal() * 0.01
    Ht = np.array([[s * np.cos(th), -s * np.sin(th), rng.normal() * 8], [s * np.sin(th), s * np.cos(th), rng.normal() * 8],
                   [rng.normal() * 1e-5, rng.normal() * 1e-5, 1]])
    p = np.c_[a, np.ones(n)] @ Ht.T
    b = (p[:, :2] / p[:, 2:] + rng.normal(size=(n, 2)) * noise).astype(np.float32)
    o = rng.random(n) < out_frac
    b[o] = (rng.random((int(o.sum()), 2)) * [1920, 1080]).astype(np.float32)
    return a, b


def _pack(sets, dev):
    cnt = np.array([len(a) for a, _ in sets], np.int32)
    cap = (cnt + 3) // 4 * 4 + 4
    off = np.zeros(len(sets), np.int32)
    off[1:] = np.cumsum(cap)[:-1]
    pts = np.zeros((int(cap.sum()), 4), np.float32)
    for i, (a, b) in enumerate(sets):
        pts[off[i]:off[i] + cnt[i], :2] = a
        pts[off[i]:off[i] + cnt[i], 2:] = b
    return (torch.from_numpy(pts).to(dev), torch.from_numpy(off).to(dev), torch.from_numpy(cnt).to(dev), off, cnt)


def _px(H, a):
    p = np.c_[a.astype(np.float64), np.ones(len(a))] @ np.asarray(H, np.float64).reshape(3, 3).T
    return p[:, :2] / p[:, 2:]


def test_find_homography_vs_oracle(engine):
    rng = np.random.default_rng(11)
    sets = [_mk(rng, n, f) for n, f in [(1200, 0.2), (700, 0.5), (300, 0.0), (64, 0.3), (5, 0.0), (4, 0.0), (3, 0.0),
                                         (2000, 0.8), (9, 0.4)]]
    pts, off, cnt, off_h, cnt_h = _pack(sets, engine.device)
    status = torch.zeros(len(sets), dtype=torch.int32, device=engine.device)
    n_hyp, seed, base, level = 1024, 7, 100, 1
    out = engine.find_homography(pts, off, cnt, status, int(cnt_h.max()), n_hyp, seed, base, level, 3.0, 0.0, 4)
    torch.cuda.synchronize()
    st = status.cpu().numpy()
    for i, (a, b) in enumerate(sets):
        ref = ransac.find_homography_seeded(a, b, n_hyp, seed, base + i, level, 3.0)
        assert st[i] == ref["status"], (i, st[i], ref["status"])
        if ref["status"] != 0:
            continue
        o, m = off_h[i], cnt_h[i]
        if ref["hyp"] is not None:
            assert int(out["best_hyp"][i]) == ref["hyp"]["best"], i                       # same winning hypothesis
            assert int(out["best_cnt"][i]) == ref["hyp"]["best_count"], i
            Hb = out["H_best"][i].cpu().numpy()
            assert np.array_equal(Hb, ref["hyp"]["H_all"][ref["hyp"]["best"]]), i         # bit-exact f64 4-point solve
        mb = out["mask_best"][o:o + m].cpu().numpy().astype(bool)
        assert np.array_equal(mb, ref["mask_best"]), i                                    # criterion (b): bit-exact masks
        H = out["H"][i].cpu().numpy()
        inl = ref["mask_best"]
        err = np.linalg.norm(_px(H, a[inl]) - _px(ref["H"], a[inl]), axis=1).mean()
        assert err < 1e-3, (i, err)                                                       # criterion (c)
        # final mask: identical except where the f32 error sits on the threshold
        gm = out["mask"][o:o + m].cpu().numpy()
        e = ransac.reproj_err32(ref["H"].ravel(), a, b)
        diff = gm != ref["mask"]
        assert (np.abs(e[diff] - 9.0) < 1e-2).all(), i
        assert int(out["inl_cnt"][i]) == int(gm.sum())


def test_find_homography_against_opencv_refit(engine):
    """criterion (c) against live OpenCV: cv2.findHomography(pts[mask_best], method 0)."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(12)
    sets = [_mk(rng, n, f) for n, f in [(1500, 0.3), (400, 0.1), (90, 0.5)]]
    pts, off, cnt, off_h, cnt_h = _pack(sets, engine.device)
    status = torch.zeros(len(sets), dtype=torch.int32, device=engine.device)
    out = engine.find_homography(pts, off, cnt, status, int(cnt_h.max()), 1024, 1, 0, 1)
    for i, (a, b) in enumerate(sets):
        o, m = off_h[i], cnt_h[i]
        mb = out["mask_best"][o:o + m].cpu().numpy().astype(bool)
        Hcv, _ = cv2.findHomography(a[mb], b[mb], 0)
        err = np.linalg.norm(_px(out["H"][i].cpu().numpy(), a[mb]) - _px(Hcv, a[mb]), axis=1).mean()
        assert err < 1e-3, (i, err)


def test_pre_transform_and_gate(engine):
    rng = np.random.default_rng(13)
    a, b = _mk(rng, 500, 0.5)
    T = np.array([[1.01, 0.02, 5.0], [-0.01, 0.99, -3.0], [1e-5, -2e-5, 1.0]])
    pts, off, cnt, off_h, cnt_h = _pack([(a, b)], engine.device)
    status = torch.zeros(1, dtype=torch.int32, device=engine.device)
    pre = torch.from_numpy(T.reshape(1, 9)).to(engine.device)
    out = engine.find_homography(pts, off, cnt, status, 500, 512, 3, 5, 2, 3.0, 0.7, 5, pre_H=pre)
    def tr(p):
        q = np.c_[p.astype(np.float64), np.ones(len(p))]
        X = (T[0, 0] * q[:, 0] + T[0, 1] * q[:, 1]) + T[0, 2]
        Y = (T[1, 0] * q[:, 0] + T[1, 1] * q[:, 1]) + T[1, 2]
        W = (T[2, 0] * q[:, 0] + T[2, 1] * q[:, 1]) + T[2, 2]
        return np.stack([X / W, Y / W], 1).astype(np.float32)
    ref = ransac.find_homography_seeded(tr(a), tr(b), 512, 3, 5, 2, 3.0)
    assert np.array_equal(out["mask_best"][:500].cpu().numpy().astype(bool), ref["mask_best"])
    assert int(status[0]) == 6          # ~50 % inliers < 70 % -> HomographyException


def test_static_filter_vs_oracle(engine, golden):
    sets, Hs = [], []
    for i in range(int(golden["fh_n"])):
        sets.append((golden[f"fh{i}_a"], golden[f"fh{i}_b2"]))
        Hs.append(golden[f"fh{i}_Hr"].ravel())
    rng = np.random.default_rng(5)
    a, b = _mk(rng, 800, 0.3, noise=2.0)
    sets.append((a, b)); Hs.append(np.eye(3).ravel())
    pts, off, cnt, off_h, cnt_h = _pack(sets, engine.device)
    status = torch.zeros(len(sets), dtype=torch.int32, device=engine.device)
    H = torch.from_numpy(np.array(Hs)).to(engine.device)
    out_pts, out_cnt, best_r, flags = engine.static_filter(pts, off, cnt, H, status)
    for i, (a, b) in enumerate(sets):
        keep, br, bad = static_filter.static_points(Hs[i], a, b)
        assert int(best_r[i]) == br and int(out_cnt[i]) == len(keep) and bool(int(flags[i])) == bad
        g = out_pts[off_h[i]:off_h[i] + len(keep)].cpu().numpy()
        assert np.array_equal(g[:, :2], a[keep]) and np.array_equal(g[:, 2:], b[keep])
    for i in range(int(golden["fh_n"])):      # and against the reference's own output
        g = out_pts[off_h[i]:off_h[i] + int(out_cnt[i])].cpu().numpy()
        assert np.array_equal(g[:, :2], golden[f"fh{i}_static_a"]) and np.array_equal(g[:, 2:], golden[f"fh{i}_static_b"])


@pytest.mark.parametrize("scale,persp", [(1.0, 1e-5), (8.0, 1e-5), (1.0, 2e-4), (0.05, 1e-5)])
def test_fused_fast_path_equals_exact_scoring(engine, scale, persp):
    """The FFMA + rcp.approx classification must give the same inlier counts as the exactly
    rounded formula everywhere (EVZ_OPT_RANSAC_EXACT forces the latter)."""
    rng = np.random.default_rng(int(scale * 100) + int(persp * 1e6))
    sets = []
    for k in range(24):
        n = int(rng.integers(200, 2000))
        a = (rng.random((n, 2)) * [1920, 1080] * scale).astype(np.float32)
        th = rng.normal() * 0.02
        Ht = np.array([[np.cos(th), -np.sin(th), rng.normal() * 8 * scale], [np.sin(th), np.cos(th), rng.normal() * 8 * scale],
                       [rng.normal() * persp / scale, rng.normal() * persp / scale, 1]])
        p = np.c_[a, np.ones(n)] @ Ht.T
        # residuals spread densely around the 3 px threshold to exercise the band
        b = (p[:, :2] / p[:, 2:] + rng.normal(size=(n, 2)) * rng.choice([0.3, 1.5, 2.2], size=(n, 1))).astype(np.float32)
        o = rng.random(n) < 0.3
        b[o] = (rng.random((int(o.sum()), 2)) * [1920, 1080] * scale).astype(np.float32)
        sets.append((a, b))
    pts, off, cnt, off_h, cnt_h = _pack(sets, engine.device)
    outs = []
    for exact in (0, 1):
        engine.set_option(1, exact)
        status = torch.zeros(len(sets), dtype=torch.int32, device=engine.device)
        outs.append((engine.find_homography(pts, off, cnt, status, int(cnt_h.max()), 2048, 3, 50, 1), status))
    engine.set_option(1, 0)
    (f, sf), (e, se) = outs
    assert torch.equal(sf, se)
    for k in ("best_hyp", "best_cnt", "mask_best", "H_best", "mask", "inl_cnt"):
        assert torch.equal(f[k], e[k]), k
    i = 3
    ref = ransac.find_homography_seeded(sets[i][0], sets[i][1], 2048, 3, 50 + i, 1)
    assert int(f["best_hyp"][i]) == ref["hyp"]["best"] and int(f["best_cnt"][i]) == ref["hyp"]["best_count"]


@pytest.mark.parametrize("level", [1, 2])
def test_exact_pruning_equals_scoring_every_hypothesis(engine, level):
    """Once a hypothesis counts every point as a sure inlier, hypotheses with a larger index cannot win the
    (count desc, index asc) arg-max and are skipped (level 2 additionally probes the first 64).  The result
    must equal scoring all hypotheses (EVZ_OPT_RANSAC_NO_PRUNE) and the oracle."""
    rng = np.random.default_rng(77 + level)
    sets = []
    for k in range(40):
        n = int(rng.integers(5, 900))
        a = (rng.random((n, 2)) * [1920, 1080]).astype(np.float32)
        th = rng.normal() * 0.01
        Ht = np.array([[np.cos(th), -np.sin(th), rng.normal() * 5], [np.sin(th), np.cos(th), rng.normal() * 5],
                       [rng.normal() * 1e-6, rng.normal() * 1e-6, 1]])
        p = np.c_[a, np.ones(n)] @ Ht.T
        noise = [0.0, 0.05, 0.3, 1.2][k % 4]              # k % 4 < 3: all-inlier sets (pruning fires); else mixed
        b = (p[:, :2] / p[:, 2:] + rng.normal(size=(n, 2)) * noise).astype(np.float32)
        if k % 8 == 7:
            o = rng.random(n) < 0.3
            b[o] = (rng.random((int(o.sum()), 2)) * [1920, 1080]).astype(np.float32)
        sets.append((a, b))
    pts, off, cnt, off_h, cnt_h = _pack(sets, engine.device)
    outs = []
    for no_prune in (0, 1):
        engine.set_option(3, no_prune)
        status = torch.zeros(len(sets), dtype=torch.int32, device=engine.device)
        outs.append((engine.find_homography(pts, off, cnt, status, int(cnt_h.max()), 700, 9, 1000, level), status))
    engine.set_option(3, 0)
    (f, sf), (e, se) = outs
    assert torch.equal(sf, se)
    for k in ("best_hyp", "best_cnt", "mask_best", "H_best", "mask", "inl_cnt", "H"):
        assert torch.equal(f[k], e[k]), k
    fired = int((f["best_cnt"].cpu() == torch.from_numpy(cnt_h.astype(np.int32))).sum())
    assert fired >= 10                                       # the pruned path was actually exercised
    for i in (0, 1, 2, 7, 13):
        ref = ransac.find_homography_seeded(sets[i][0], sets[i][1], 700, 9, 1000 + i, level)
        assert int(f["best_hyp"][i]) == ref["hyp"]["best"] and int(f["best_cnt"][i]) == ref["hyp"]["best_count"]


@pytest.mark.parametrize("n,out_frac,n_hyp,level", [(4915, 0.2, 1024, 1), (4915, 0.0, 1024, 2), (1229, 0.8, 4096, 1),
                                                     (12288, 0.3, 512, 1), (2600, 0.8, 4096, 1)])
def test_find_homography_at_baseline_config_sizes(engine, n, out_frac, n_hyp, level):
    """Sizes of BASELINE configs 3 (8192 keypoints -> ~4 900 matches x 1 024 hypotheses) and 4 (~1 200 matches, 80 %
    outliers, 4 096 hypotheses), and the largest supported point set: winning hypothesis, its count, its f64 model
    and its inlier mask bit-exact against the oracle (criterion b), final H within 1e-3 px (criterion c)."""
    rng = np.random.default_rng(n + n_hyp + level)
    sets = [_mk(rng, n, out_frac), _mk(rng, n - 37, out_frac)]
    pts, off, cnt, off_h, cnt_h = _pack(sets, engine.device)
    status = torch.zeros(len(sets), dtype=torch.int32, device=engine.device)
    out = engine.find_homography(pts, off, cnt, status, int(cnt_h.max()), n_hyp, 5, 300, level, 3.0, 0.0, 4)
    st = status.cpu().numpy()
    for i, (a, b) in enumerate(sets):
        ref = ransac.find_homography_seeded(a, b, n_hyp, 5, 300 + i, level, 3.0)
        assert st[i] == ref["status"], (i, st[i], ref["status"])
        if ref["status"] != 0:
            continue
        o, m = off_h[i], cnt_h[i]
        assert int(out["best_hyp"][i]) == ref["hyp"]["best"] and int(out["best_cnt"][i]) == ref["hyp"]["best_count"], i
        assert np.array_equal(out["H_best"][i].cpu().numpy(), ref["hyp"]["H_all"][ref["hyp"]["best"]]), i
        assert np.array_equal(out["mask_best"][o:o + m].cpu().numpy().astype(bool), ref["mask_best"]), i
        inl = ref["mask_best"]
        err = np.linalg.norm(_px(out["H"][i].cpu().numpy(), a[inl]) - _px(ref["H"], a[inl]), axis=1).mean()
        assert err < 1e-3, (i, err)


def test_find_homography_degenerate_sets(engine):
    """Inputs on which most or all hypotheses are rejected or the fit is exact: collinear points (no valid sample: the
    reference's findHomography returns None -> "can't find homography matrix"), every point repeated three times, an exact
    translation without noise (all residuals zero: the LM loop must not start), a set whose inliers lie on a line (the
    winning samples are near-degenerate), coordinates a hundred times larger.  Status, winner, masks against the oracle."""
    rng = np.random.default_rng(23)
    n = 400
    x = rng.permutation(1900)[:n].astype(np.float32)                                      # integers: exactly collinear in f32
    col_a = np.c_[x, 0.5 * x + 10].astype(np.float32)
    col = (col_a, (col_a + np.float32([3, -2])).astype(np.float32))                       # all collinear
    a3, b3 = _mk(rng, 150, 0.2)
    rep = (np.repeat(a3, 3, 0), np.repeat(b3, 3, 0))                                      # every correspondence three times
    at = (rng.random((300, 2)) * [1920, 1080]).astype(np.float32)
    at = np.round(at)                                                                     # exact in f32 after the shift
    tr = (at, (at + np.float32([5, -7])).astype(np.float32))                              # exact translation, zero residual
    al, bl = _mk(rng, 500, 0.0)
    line = rng.random(500) < 0.8
    al[line, 1] = (0.3 * al[line, 0] + 100).astype(np.float32)                            # 80 % of the points on one line ...
    bl[line] = (al[line] + np.float32([4, 4])).astype(np.float32)
    bl[~line] = (al[~line] + np.float32([4, 4]) + rng.normal(size=(int((~line).sum()), 2)) * 0.3).astype(np.float32)
    big_a, big_b = _mk(rng, 600, 0.3)
    big = ((big_a * 100).astype(np.float32), (big_b * 100).astype(np.float32))            # 192 000 px wide: no inlier at 3 px but the exact ones
    sets = [col, rep, tr, (al, bl), big]
    pts, off, cnt, off_h, cnt_h = _pack(sets, engine.device)
    status = torch.zeros(len(sets), dtype=torch.int32, device=engine.device)
    n_hyp, seed, base, level = 512, 3, 7, 1
    out = engine.find_homography(pts, off, cnt, status, int(cnt_h.max()), n_hyp, seed, base, level, 3.0, 0.0, 4)
    torch.cuda.synchronize()
    st = status.cpu().numpy()
    for i, (a, b) in enumerate(sets):
        ref = ransac.find_homography_seeded(a, b, n_hyp, seed, base + i, level, 3.0)
        assert st[i] == ref["status"], (i, st[i], ref["status"])
        if ref["status"] != 0:
            continue
        o, m = off_h[i], cnt_h[i]
        assert int(out["best_hyp"][i]) == ref["hyp"]["best"] and int(out["best_cnt"][i]) == ref["hyp"]["best_count"], i
        assert np.array_equal(out["mask_best"][o:o + m].cpu().numpy().astype(bool), ref["mask_best"]), i
        inl = ref["mask_best"]
        H = out["H"][i].cpu().numpy()
        assert np.isfinite(H).all(), i
        scale = 100.0 if i == 4 else 1.0
        err = np.linalg.norm(_px(H, a[inl]) - _px(ref["H"], a[inl]), axis=1).mean()
        assert err < 1e-3 * scale, (i, err)
    assert st[0] != 0                                                                     # collinear: no model
    assert st[2] == 0 and int(out["inl_cnt"][2]) == 300                                   # exact translation: every point


def test_static_filter_rounding_ties_windows(engine):
    """find_point_displacement / get_largest_group_points corner cases (reference utils.py:258-325): displacements exactly on
    .5 (Python round = half to even), two bins with the same count (the first-inserted key wins), bins thousands of pixels
    apart (several histogram windows), and a displacement beyond EVZ_R_MAX (the flag, where the reference's dict would simply
    hold one more key).  Against the oracle's restatement, which is pinned to the reference's own output."""
    H = np.eye(3).ravel()
    def shifted(n, dx, start):
        a = np.c_[np.arange(start, start + n), np.full(n, 7.0)].astype(np.float32)
        return a, (a + np.float32([dx, 0])).astype(np.float32)
    # (1) 2.5 -> 2 and 3.5 -> 4 and 2.0 -> 2: bin 2 collects 5 + 4 points, bin 4 collects 6
    parts = [shifted(5, 2.5, 0), shifted(6, 3.5, 100), shifted(4, 2.0, 200)]
    s1 = (np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]))
    # (2) a tie: bins 9 (inserted first), 3 and 5 hold four points each
    parts = [shifted(4, 9.0, 0), shifted(4, 3.0, 50), shifted(4, 5.0, 90), shifted(2, 1.0, 130)]
    s2 = (np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]))
    # (3) bins 0, 5000 and 16000: three windows of the histogram, the largest group in the last one
    parts = [shifted(30, 0.0, 0), shifted(20, 5000.0, 100), shifted(40, 16000.0, 200)]
    s3 = (np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]))
    # (4) one displacement beyond EVZ_R_MAX = 16383
    parts = [shifted(10, 4.0, 0), shifted(1, 20000.0, 50)]
    s4 = (np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]))
    sets = [s1, s2, s3, s4]
    pts, off, cnt, off_h, cnt_h = _pack(sets, engine.device)
    status = torch.zeros(len(sets), dtype=torch.int32, device=engine.device)
    Hd = torch.from_numpy(np.tile(H, (len(sets), 1))).to(engine.device)
    out_pts, out_cnt, best_r, flags = engine.static_filter(pts, off, cnt, Hd, status)
    torch.cuda.synchronize()
    want_r = [2, 9, 16000, 4]
    for i, (a, b) in enumerate(sets):
        keep, br, bad = static_filter.static_points(H, a, b)
        assert br == want_r[i], (i, br)
        assert int(best_r[i]) == br and int(out_cnt[i]) == len(keep) and bool(int(flags[i])) == bad, i
        g = out_pts[off_h[i]:off_h[i] + len(keep)].cpu().numpy()
        assert np.array_equal(g[:, :2], a[keep]) and np.array_equal(g[:, 2:], b[keep]), i
    assert int(out_cnt[0]) == 9 and int(flags[3]) == 1 and int(flags[2]) == 0


@pytest.mark.parametrize("thresh", [0.75, 1.0, 5.0, 12.5])
def test_find_homography_other_thresholds(engine, thresh):
    """match_static_kps(reproj_thresh=...) (matching.py:131, 156-157): the fused fast path derives its bands from the
    threshold, so winner and masks must equal the oracle at other thresholds than the reference's default 3.0."""
    rng = np.random.default_rng(int(thresh * 100))
    sets = [_mk(rng, 900, 0.3, noise=0.8), _mk(rng, 333, 0.6, noise=1.5), _mk(rng, 1500, 0.1, noise=0.3)]
    pts, off, cnt, off_h, cnt_h = _pack(sets, engine.device)
    status = torch.zeros(len(sets), dtype=torch.int32, device=engine.device)
    out = engine.find_homography(pts, off, cnt, status, int(cnt_h.max()), 768, 5, 11, 1, thresh, 0.0, 4)
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(sets):
        ref = ransac.find_homography_seeded(a, b, 768, 5, 11 + i, 1, thresh)
        assert int(status[i]) == ref["status"] == 0, i
        o, m = off_h[i], cnt_h[i]
        assert int(out["best_hyp"][i]) == ref["hyp"]["best"] and int(out["best_cnt"][i]) == ref["hyp"]["best_count"], i
        assert np.array_equal(out["mask_best"][o:o + m].cpu().numpy().astype(bool), ref["mask_best"]), i
        inl = ref["mask_best"]
        err = np.linalg.norm(_px(out["H"][i].cpu().numpy(), a[inl]) - _px(ref["H"], a[inl]), axis=1).mean()
        assert err < 1e-3, (i, err)
