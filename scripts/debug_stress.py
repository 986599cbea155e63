"""Debug helper: the ragged stress launch per variant with a sync after each call."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import evenvizion_b200 as evz
eng = evz.GeometryEngine(0)
rng = np.random.default_rng(123)
sizes = [0, 1, 2, 7, 8, 9, 127, 128, 129, 255, 256, 257, 511, 512, 513, 700, 1023, 1024, 1025, 1500, 2047, 2048, 2049, 2600]
counts = [int(rng.choice(sizes)) for _ in range(90)]
counts[5], counts[17] = 127, 2049
tot = sum(counts)
base = rng.integers(0, 256, (3000, 128)).astype(np.uint8)
desc = np.empty((tot, 128), np.uint8)
o = 0
for n in counts:
    sel = rng.integers(0, len(base), n)
    d = base[sel].astype(np.int32) + (rng.integers(-1, 2, (n, 128)) * (rng.random((n, 1)) < 0.5))
    desc[o:o + n] = np.clip(d, 0, 255).astype(np.uint8)
    o += n
coords = rng.random((tot, 2)).astype(np.float32)
st = eng.ingest(desc, coords, counts)
F = len(counts)
variants = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [5, 0]
lim = int(sys.argv[2]) if len(sys.argv) > 2 else F - 1
for v in variants:
    eng.set_option(2, v)
    for hi in (1, 2, 4, 8, 16, 32, 64, lim):
        hi = min(hi, lim)
        print("variant", v, "pairs", hi, "counts", counts[:hi + 1] if hi <= 8 else "...", flush=True)
        r = eng.match(st, list(range(1, hi + 1)), list(range(0, hi)))
        torch.cuda.synchronize()
        print("  ok", flush=True)
print("---- test pattern", flush=True)
for pq, pt in ((list(range(1, F)), list(range(0, F - 1))), ([5, 0, 17, 33, 60], [5, 40, 17, 2, 88])):
    for v in (0, 1, 2, 3, 4, 5):
        eng.set_option(2, v)
        print("variant", v, "pairs", len(pq), flush=True)
        outs = [eng.match(st, pq, pt) for _ in range(2 if v == 0 else 1)]
        torch.cuda.synchronize()
        print("  ok", flush=True)
