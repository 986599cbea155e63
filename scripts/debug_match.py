"""First-light diagnostics for the tcgen05 match kernel: tiny shapes, mismatch statistics."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import evenvizion_b200 as evz
from oracle import matching

eng = evz.GeometryEngine(0)
print("sm_count", eng.sm_count, torch.cuda.get_device_name(0), flush=True)
rng = np.random.default_rng(0)


def run(nq, nt, tag, pattern="rand"):
    if pattern == "rand":
        q = rng.integers(0, 256, (nq, 128)).astype(np.uint8)
        t = rng.integers(0, 256, (nt, 128)).astype(np.uint8)
    elif pattern == "onehot":       # q row i has a single 1 at byte (i % 128); t row j = j % 251 everywhere
        q = np.zeros((nq, 128), np.uint8); q[np.arange(nq), np.arange(nq) % 128] = 1
        t = np.tile((np.arange(nt) % 251).astype(np.uint8)[:, None], (1, 128))
        t[:, :] = (np.arange(nt)[:, None] * 7 + np.arange(128)[None, :] * 13) % 256
    frames_d = np.concatenate([t, q]); frames_c = np.zeros((nt + nq, 2), np.float32)
    st = eng.ingest(frames_d, frames_c, [nt, nq])
    r = eng.match(st, [1], [0])
    torch.cuda.synchronize()
    o = int(st.row_off_h[1])
    gi = r.top2_idx[o:o + nq].cpu().numpy(); gd = r.top2_d2[o:o + nq].cpu().numpy().astype(np.int64)
    idx, d2 = matching.knn_top2(q, t)
    bad_i = (gi != idx).any(1); bad_d = (gd != d2).any(1)
    print(f"[{tag}] nq={nq} nt={nt} {pattern}: idx mismatch rows {bad_i.sum()}, d2 mismatch rows {bad_d.sum()}", flush=True)
    if bad_i.any() or bad_d.any():
        rows = np.nonzero(bad_i | bad_d)[0]
        print("   first bad rows", rows[:16], "last", rows[-4:])
        for rr in rows[:4]:
            print("   row", rr, "gpu", gi[rr], gd[rr], "ref", idx[rr], d2[rr])
        # is the GPU value a valid distance to *some* column?  (layout / swizzle diagnosis)
        full = ((q[rows[:4], None, :].astype(np.int64) - t[None, :, :].astype(np.int64)) ** 2).sum(-1)
        for k, rr in enumerate(rows[:4]):
            hit = np.nonzero(full[k] == gd[rr, 0])[0]
            print("   row", rr, "gpu d2[0] equals true distance to columns", hit[:8])
    return not (bad_i.any() or bad_d.any())


ok = True
for args in [(128, 256, "1sub-1tile"), (256, 256, "2sub-1tile"), (256, 512, "2sub-2tile"), (100, 70, "ragged"),
             (128, 256, "onehot", "onehot"), (2048, 2048, "full"), (1, 5, "tiny"), (300, 1, "nt1")]:
    try:
        ok &= run(*args[:3], *(args[3:]))
    except Exception as e:
        print("EXC", args, e, flush=True); ok = False
print("ALL OK" if ok else "FAILURES")
