"""Micro-benchmark of the post-processing kernels at config-5 scale (BASELINE.json configs[4]):
None-H fill + cumulative 3x3 prefix product over P pairs, object-coordinate remap of k points per
frame, dense max-movement.  Prints achieved HBM GB/s against the algorithmic bytes of DESIGN.md."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import evenvizion_b200 as evz

P = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 8
eng = evz.GeometryEngine(0)
dev = eng.device
rng = np.random.default_rng(0)
G = np.tile(np.eye(3), (P, 1, 1))
G[:, 0, 2] = rng.normal(0, 2, P); G[:, 1, 2] = rng.normal(0, 2, P)
G[:, 0, 1] = rng.normal(0, 1e-3, P); G[:, 1, 0] = -G[:, 0, 1]; G[:, 2, 0] = rng.normal(0, 1e-7, P)
Gd = torch.from_numpy(G.reshape(P, 9)).to(dev)
status = torch.from_numpy((rng.random(P) < 0.02).astype(np.int32)).to(dev)


def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n


ms = timeit(lambda: eng.chain_scan(Gd, status, True))
S, Hf, _ = eng.chain_scan(Gd, status, True)
print(f"chain_scan   P={P}: {ms*1e3:8.1f} us   {(72 + 4 + 72 + 72) * P / ms / 1e6:8.1f} GB/s (72 B G + 4 B status in, 72 B S + 72 B H_fixed out per pair)")
n = (P + 1) * K
pts = torch.from_numpy(rng.random((n, 2)) * [1170, 658]).to(dev)
fidx = torch.arange(P + 1, dtype=torch.int32, device=dev).repeat_interleave(K).contiguous()
Sall = torch.cat([torch.eye(3, dtype=torch.float64, device=dev).reshape(1, 9), S]).contiguous()
for inv in (False, True):
    ms = timeit(lambda: eng.remap(pts, fidx, Sall, 400 / 1170, 224 / 658, inv))
    print(f"remap inv={int(inv)} n={n}: {ms*1e3:8.1f} us   {(16 + 4 + 16) * n / ms / 1e6:8.1f} GB/s (16 B in + 4 B frame id + 16 B out per point; S rows hit L2)")
F = min(P, 2000)
ms = timeit(lambda: eng.max_movement(Sall, F, 224, 400), 5)
print(f"max_movement {F} frames x 224x400: {ms*1e3:8.1f} us   {F*224*400/ms/1e6:8.1f} Gpixel/s")
