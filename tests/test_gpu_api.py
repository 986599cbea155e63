"""The reference-facing Python API (evenvizion_b200.processing) on the GPU.  Needs a B200."""
import json

import numpy as np
import pytest

from oracle import chain, pipeline, ransac, static_filter

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def proc(engine):
    import evenvizion_b200
    evenvizion_b200._default_engine = engine
    import evenvizion_b200.processing as p
    return p


def test_match_kps_equals_reference_output(proc, golden):
    for i in range(int(golden["mk_n"])):
        kq = proc.KeyPoints(golden[f"mk{i}_qc"], golden[f"mk{i}_qd"].astype(np.float32))   # SIFT-style f32 input
        kt = proc.KeyPoints(golden[f"mk{i}_tc"], golden[f"mk{i}_td"].astype(np.float32))
        pa, pb = kq.match_kps(kt)
        assert isinstance(pa, list) and pa[0].dtype == np.float32 and pa[0].shape == (2,)
        assert np.array_equal(np.array(pa), golden[f"mk{i}_pts_a"]) and np.array_equal(np.array(pb), golden[f"mk{i}_pts_b"])


def test_match_static_kps_and_compute_homography(proc, golden):
    for i in range(int(golden["mk_n"])):
        qc, qd, tc, td = (golden[f"mk{i}_{k}"] for k in ("qc", "qd", "tc", "td"))
        sa, sb = proc.KeyPoints(qc, qd).match_static_kps(proc.KeyPoints(tc, td), n_hyp=1024, seed=0, pair_id=i)
        rp = pipeline.pair_geometry(qc, qd, tc, td, n_hyp=1024, seed=0, pair_id=i)
        assert np.array_equal(sa, rp["static_a"]) and np.array_equal(sb, rp["static_b"])
        assert isinstance(sa, np.ndarray) and sa.dtype == np.float32
        H = proc.compute_homography(sa, sb, n_hyp=1024, seed=0, pair_id=i)
        assert H.shape == (3, 3) and H.dtype == np.float64
        p = np.c_[sa.astype(np.float64), np.ones(len(sa))]
        a = p @ H.T; b = p @ rp["H"].T
        assert np.abs(a[:, :2] / a[:, 2:] - b[:, :2] / b[:, 2:]).mean() < 1e-3
        # against the reference's own cv2-sampled result: agreement, not identity
        c = p @ golden[f"mk{i}_H"].T
        assert np.abs(a[:, :2] / a[:, 2:] - c[:, :2] / c[:, 2:]).mean() < 0.5


def test_exceptions_mirror_reference(proc):
    rng = np.random.default_rng(0)
    q = proc.KeyPoints(rng.random((50, 2)).astype(np.float32), rng.integers(0, 255, (50, 128)).astype(np.uint8))
    t = proc.KeyPoints(rng.random((60, 2)).astype(np.float32), rng.integers(0, 255, (60, 128)).astype(np.uint8))
    with pytest.raises(proc.NoMatchesException) as e:
        q.match_kps(t)                                    # random descriptors: nothing survives the ratio test
    assert "min_matching_pts 4" in str(e.value) and str(e.value).endswith("-> couldn't process")
    with pytest.raises(proc.NoMatchesException):
        proc.KeyPoints(q.coordinates, None).match_kps(t)
    a = (rng.random((200, 2)) * 1000).astype(np.float32)
    b = a.copy(); b[:120] = (rng.random((120, 2)) * 1000).astype(np.float32)      # 40 % inliers < 70 %
    with pytest.raises(proc.HomographyException) as e:
        proc.compute_homography(a, b)
    assert "not enough points" in str(e.value)
    with pytest.raises(ValueError):
        proc.KeyPoints(q.coordinates, rng.random((50, 128)).astype(np.float32)).match_kps(t)   # SURF-like floats


def test_find_point_displacement_groups(proc, golden):
    a, b, H = golden["fh2_a"], golden["fh2_b2"], golden["fh2_Hr"]
    g = proc.find_point_displacement(H, a, b)
    r, _ = static_filter.displacement_bins(H, a, b)
    exp = {}
    for i, v in enumerate(r.tolist()):
        exp.setdefault(v, []).append(i)
    assert g == exp and list(g) == list(exp)              # same groups, same insertion order
    ga, gb = proc.get_largest_group_points(g, a, b)
    assert np.array_equal(ga, golden["fh2_static_a"]) and np.array_equal(gb, golden["fh2_static_b"])


def test_superposition_dict_and_remap_json_roundtrip(proc, bundled, tmp_path):
    path = tmp_path / "dict_with_homography_matrix.json"
    d = dict(bundled["homography_dict"]); d["resize_info"] = bundled["resize_info"]
    path.write_text(json.dumps(d))
    hd, ri = proc.read_homography_dict(str(path))
    assert list(hd) == list(range(2, 122)) and ri == bundled["resize_info"]
    sup = proc.superposition_dict(hd)
    assert list(sup) == list(range(1, 122)) and sup[1] == [[1, 0, 0], [0, 1, 0], [0, 0, 1]]
    for k, v in bundled["superposition"].items():
        assert np.abs(np.asarray(sup[int(k)], np.float64) - np.asarray(v)).max() < 1e-9
    oc = {int(k): v for k, v in bundled["original_coordinates"].items()}
    fixed = proc.from_original_to_fix(oc, sup, bundled["original_shape"], [ri["h"], ri["w"]])
    n = bad = 0
    for k, rects in bundled["fixed_coordinates"].items():
        for r0, r1 in zip(rects, fixed[int(k)]):
            assert abs(r0["x1"] - r1["x1"]) <= 0.0100001 and abs(r0["y1"] - r1["y1"]) <= 0.0100001
            n += 1; bad += (r0["x1"] != r1["x1"]) or (r0["y1"] != r1["y1"])
    assert bad <= n // 100
    json.dumps({str(k): v for k, v in fixed.items()})     # the fixed-coordinate JSON is serialisable
    back = proc.from_fix_to_original(fixed, sup, bundled["original_shape"], [ri["h"], ri["w"]])
    for k, rects in bundled["back_to_original"].items():
        for r0, r1 in zip(rects, back[int(k)]):
            assert abs(r0["x1"] - r1["x1"]) <= 0.0200001 and abs(r0["y1"] - r1["y1"]) <= 0.0200001
    # H None entries carry the superposition (utils.py:139,145)
    hd2 = dict(hd); hd2[50] = {"H": None}
    sup2 = proc.superposition_dict(hd2)
    ref2 = chain.superposition_dict(hd2)
    for k in (49, 50, 51, 121):
        assert np.abs(np.asarray(sup2[k], np.float64) - np.asarray(ref2[k], np.float64)).max() < 1e-9


@pytest.mark.parametrize("mode", ["reference", "parallel"])
def test_video_driver_on_bundled_clip_features(proc, golden, mode):
    n = int(golden["clip_n"])
    feats = {"SIFT": [(golden[f"clip{f}_c"], golden[f"clip{f}_d"]) for f in range(n)]}
    H_list, status = proc.geometry_from_features(feats, True, mode, n_hyp=1024, seed=0)
    ref = pipeline.video_chain(feats["SIFT"], n_hyp=1024, seed=0, reference_exact=(mode == "reference"))
    assert np.array_equal(status, ref["status"])
    grid = np.stack(np.meshgrid(np.linspace(0, 400, 9), np.linspace(0, 224, 6)), -1).reshape(-1, 2)
    hd = {k + 2: {"H": np.asarray(H).tolist()} for k, H in enumerate(H_list)}
    sup = chain.superposition_dict(hd)                      # the reference's own left fold over our output
    for k in range(n - 1):
        p = np.c_[grid, np.ones(len(grid))]
        a = p @ np.asarray(sup[k + 2], np.float64).T; b = p @ ref["S"][k + 1].T
        assert np.abs(a[:, :2] / a[:, 2:] - b[:, :2] / b[:, 2:]).max() < 1e-2, (mode, k)    # criterion (d)


def test_multi_feature_concat_dedup(proc, golden, engine):
    import torch
    from evenvizion_b200.processing.video_processing import _dedup_concat
    from oracle import matching
    rng = np.random.default_rng(4)
    a1 = rng.integers(0, 5, (40, 2)).astype(np.float32); b1 = rng.random((40, 2)).astype(np.float32)
    a2 = rng.integers(0, 5, (30, 2)).astype(np.float32); b2 = rng.random((30, 2)).astype(np.float32)
    t = lambda a, b: torch.from_numpy(np.c_[a, b]).to(engine.device)
    out = _dedup_concat([t(a1, b1), t(a2, b2)]).cpu().numpy()
    na, nb, _, _ = matching.remove_double_matching(np.r_[a1, a2], np.r_[b1, b2])
    assert np.array_equal(out[:, :2], na) and np.array_equal(out[:, 2:], nb)


def test_video_geometry_streamed_equals_single_batch(engine):
    """Chunked H2D/compute overlap (host input) must give bit-identical results to one batch, for
    ragged frames, empty frames and chunk sizes that do not divide the frame count."""
    import torch
    from evenvizion_b200 import synth
    ch = synth.make_chain(23, 700, seed=11, device="cpu", unmatched_frac=0.1)
    counts = [700] * 23
    counts[5], counts[9], counts[17] = 300, 0, 513
    desc = torch.cat([ch["desc"][f, :n] for f, n in enumerate(counts)])
    coords = torch.cat([ch["coords"][f, :n] for f, n in enumerate(counts)])
    ref = engine.video_geometry(desc.cuda(), coords.cuda(), counts, n_hyp=256, seed=5, pair_id_base=40)   # single batch
    for chunk in (4, 7, 22):
        out = engine.video_geometry(desc.pin_memory(), coords.pin_memory(), counts, n_hyp=256, seed=5, pair_id_base=40,
                                    chunk_frames=chunk)
        for k in ("status", "G", "S", "H_fixed"):
            assert np.array_equal(out[k], ref[k]), (chunk, k)
        a, b = out["results"], ref["results"]
        ro = out["store"].row_off_h
        valid = np.concatenate([np.arange(ro[f], ro[f] + counts[f]) for f in range(1, 23)])     # written rows only
        valid = torch.from_numpy(valid).cuda()
        assert torch.equal(a.top2_idx[valid], b.top2_idx[valid]) and torch.equal(a.top2_d2[valid], b.top2_d2[valid])
        assert torch.equal(a.m_cnt, b.m_cnt)
        assert torch.equal(a.static_cnt, b.static_cnt) and torch.equal(a.best_hyp2, b.best_hyp2)
