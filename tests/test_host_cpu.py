"""CPU-only tests: the C ABI library loads and exports what include/evz.h declares, the product
fails loudly without a GPU, and the host logic of the multi-GPU path (gloo, world size 2)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import evenvizion_b200
from evenvizion_b200 import _lib
from evenvizion_b200.distributed import shard_range, seeds_from_summaries
from evenvizion_b200.engine import GeometryEngine
from oracle import chain

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "evz.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(evz_\w+)\s*\(([^;{]*?)\)\s*;", text):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def test_library_exports_every_declared_symbol():
    decl = _declared()
    assert len(decl) >= 13
    lib = _lib.load()
    for name, nargs in decl.items():
        assert hasattr(lib, name), f"{name} declared in evz.h but not exported by libevz.so"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
        assert len(_lib.SIGNATURES[name]) == nargs, f"{name}: evz.h has {nargs} parameters, ctypes binding {len(_lib.SIGNATURES[name])}"
    assert lib.evz_version() == 100


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.evz_create(0, ctypes.byref(h)) == -4            # EVZ_E_NODEVICE
    assert b"no CPU fallback" in lib.evz_last_error(None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        GeometryEngine(0)
    import evenvizion_b200.processing as p
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        p.KeyPoints(np.zeros((5, 2), np.float32), np.zeros((5, 128), np.uint8)).match_kps(
            p.KeyPoints(np.zeros((5, 2), np.float32), np.zeros((5, 128), np.uint8)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        p.superposition_dict({2: {"H": np.eye(3).tolist()}})


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "evenvizion_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, f)


def test_layout_and_shards():
    ro = GeometryEngine.layout([300, 1, 0, 256, 257])
    assert ro.tolist() == [0, 512, 768, 1024, 1280, 1792]
    for n, w in [(10, 4), (100000, 8), (3, 8), (0, 2)]:
        r = [shard_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
        assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def _summary(G, valid, policy):
    """Host restatement of the evz_chain_scan shard summary (include/evz.h) via the oracle."""
    n = len(G)
    idx = np.nonzero(valid)[0]
    s = np.zeros(20)
    if len(idx) == 0:
        s[0:9] = np.eye(3).ravel(); s[9:18] = np.eye(3).ravel(); s[18] = n; s[19] = 0
        return s
    lead = int(idx[0])
    Gf = chain.fill_none(G[lead:], valid[lead:], policy)
    s[0:9] = chain.chain_products(Gf)[-1].ravel()
    s[9:18] = G[idx[-1]].ravel(); s[18] = lead; s[19] = 1
    return s


def _rand_chain(P, seed):
    rng = np.random.default_rng(seed)
    G = np.tile(np.eye(3), (P, 1, 1))
    G[:, :2, :] += rng.normal(size=(P, 2, 3)) * [0.002, 0.002, 2.0]
    valid = rng.random(P) > 0.15
    valid[:2] = False
    valid[P // 2 - 3:P // 2 + 4] = False          # failures straddling the shard boundary
    return G, valid


@pytest.mark.parametrize("policy", [True, False])
def test_seeds_from_summaries_match_global_chain(policy):
    G, valid = _rand_chain(240, 1)
    ref = chain.chain_products(chain.fill_none(G, valid, policy))
    world = 5
    sums = [_summary(G[a:b], valid[a:b], policy) for a, b in (shard_range(240, r, world) for r in range(world))]
    for r in range(world):
        a, b = shard_range(240, r, world)
        S, Gs = seeds_from_summaries(sums, r, policy)
        assert np.abs(S - ref[a]).max() < 1e-9 * max(1, np.abs(ref[a]).max())
        if policy and valid[:a].any():
            assert np.array_equal(Gs, G[np.nonzero(valid[:a])[0][-1]])


def _gloo_worker(rank, world, port, policy, q):
    import torch.distributed as dist
    from evenvizion_b200.distributed import all_gather_summaries
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    G, valid = _rand_chain(120, 7)
    a, b = shard_range(120, rank, world)
    mine = torch.from_numpy(_summary(G[a:b], valid[a:b], policy))
    sums = all_gather_summaries(mine)
    S, Gs = seeds_from_summaries(sums, rank, policy)
    ref = chain.chain_products(chain.fill_none(G, valid, policy))
    # seeded local chain == slice of the global chain
    Gf = chain.fill_none(G[a:b], valid[a:b], policy)
    if policy and Gs is not None:
        lead = int(np.argmax(valid[a:b])) if valid[a:b].any() else b - a
        Gf[:lead] = Gs
    loc = chain.chain_products(Gf)
    out = np.array([(S @ loc[i]) / (S @ loc[i])[2, 2] for i in range(1, len(loc))])
    q.put((rank, float(np.abs(out - ref[a + 1:b + 1]).max())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("policy", [True, False])
def test_two_rank_gloo_all_gather_seeding(policy):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + (1 if policy else 0)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, policy, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r for r, _ in res) == [0, 1]
    assert max(e for _, e in res) < 1e-9


def test_streaming_json_writers_match_json_dump(tmp_path, bundled):
    """the streaming writers emit exactly the bytes json.dump produces for the reference's dicts,
    and the array readers invert them (incl. the bundled dict_with_homography_matrix.json)."""
    import io
    import json
    from evenvizion_b200.processing import formats
    rng = np.random.default_rng(4)
    P = 9000                                                    # more than one write chunk
    H = rng.normal(size=(P, 3, 3)) * np.array([1.0, 1e-3, 1e2])[None, :, None]
    H[:, 2, 2] = 1.0
    H[5] = [[1, 0, 0], [0, 1, 0], [0, 0, 1]]
    valid = np.ones(P, bool)
    valid[[3, 4000]] = False
    ref = {str(k + 2): {"H": (H[k].tolist() if valid[k] else None)} for k in range(P)}
    ref["resize_info"] = {"h": 224, "w": 400}
    buf = io.StringIO()
    formats.dump_homography_dict(buf, H, {"h": 224, "w": 400}, valid=valid)
    assert buf.getvalue() == json.dumps(ref)
    path = tmp_path / "dict_with_homography_matrix.json"
    formats.dump_homography_dict(path, H, {"h": 224, "w": 400}, valid=valid)
    frames, H2, v2, ri = formats.load_homography_arrays(path)
    assert np.array_equal(frames, np.arange(2, P + 2)) and np.array_equal(v2, valid) and ri == {"h": 224, "w": 400}
    assert np.array_equal(H2[valid], H[valid]) and np.isnan(H2[~valid]).all()
    with pytest.raises(ValueError):
        formats.load_homography_arrays(io.StringIO('{"2": {"H": null}}'))
    # the bundled reference file round-trips byte for byte
    hd = bundled["homography_dict"]
    keys = sorted(int(k) for k in hd)
    Hb = np.array([hd[str(k)]["H"] for k in keys])
    out = io.StringIO()
    formats.dump_homography_dict(out, Hb, bundled["resize_info"], first_frame=keys[0])
    assert out.getvalue() == json.dumps({**{str(k): hd[str(k)] for k in keys}, "resize_info": bundled["resize_info"]})
    # coordinates
    counts = rng.integers(0, 9, 500)
    xy = np.round(rng.normal(size=(int(counts.sum()), 2)) * 300, 2)
    extra = [{"x1": 0.0, "y1": 0.0, "id": int(i)} for i in range(len(xy))]
    refc, j = {}, 0
    for fr, n in enumerate(counts, start=1):
        refc[str(fr)] = []
        for _ in range(n):
            refc[str(fr)].append({"x1": float(xy[j, 0]), "y1": float(xy[j, 1]), "id": j})
            j += 1
    out = io.StringIO()
    formats.dump_coordinates(out, np.arange(1, 501), counts, xy, extra=extra)
    assert out.getvalue() == json.dumps(refc)
    fr2, c2, xy2 = formats.load_coordinates_arrays(io.StringIO(out.getvalue()))
    assert np.array_equal(fr2, np.arange(1, 501)) and np.array_equal(c2, counts) and np.array_equal(xy2, xy)


def test_bench_flann_agreement_rates():
    """bench.py's reported-only FLANN agreement (north_star): fed with the exact brute-force result it must report
    rates in [0, 1], near-perfect survivor agreement on descriptor-like data, and exactly 1.0 against itself-like
    input where FLANN is exact (tiny train sets are searched exhaustively by the KD-tree with checks >= size)."""
    import importlib.util
    from evenvizion_b200 import synth
    from oracle import matching
    spec = importlib.util.spec_from_file_location("evz_bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    ch = synth.make_chain(3, 384, seed=5, device="cpu")
    d = ch["desc"]
    gi, gs = [], []
    for k in range(2):
        idx, d2 = matching.knn_top2(d[k + 1].numpy(), d[k].numpy())
        gi.append(idx[:, 0]); gs.append(matching.ratio_survivors(idx, d2))
    out = bench.flann_agreement(d, gi, gs)
    assert out["pairs"] == 2 and 0.5 <= out["top1"] <= 1.0 and out["survivors_jaccard"] >= 0.95
    small = torch.from_numpy(np.random.default_rng(1).integers(0, 256, (2, 24, 128)).astype(np.uint8))
    idx, d2 = matching.knn_top2(small[1].numpy(), small[0].numpy())
    out = bench.flann_agreement(small, [idx[:, 0]], [matching.ratio_survivors(idx, d2)])
    assert out["top1"] == 1.0 and out["survivors_jaccard"] in (0.0, 1.0)
