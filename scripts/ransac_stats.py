"""Workload statistics of the RANSAC stages on the config-2 synthetic chain (diagnostics)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import evenvizion_b200 as evz
from evenvizion_b200 import synth
P = int(sys.argv[1]) if len(sys.argv) > 1 else 512
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
of = float(sys.argv[3]) if len(sys.argv) > 3 else 0.2
nh = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
eng = evz.GeometryEngine(0)
ch = synth.make_chain(P + 1, N, seed=0, device="cuda", outlier_frac=of)
st = eng.ingest(ch["desc"], ch["coords"])
r = eng.process_pairs(st, torch.arange(1, P + 1), torch.arange(0, P), n_hyp=nh)
torch.cuda.synchronize()
g = lambda t: t.cpu().numpy()
m, bh1, bc1, sc, bh2, bc2 = g(r.m_cnt), g(r.best_hyp1), g(r.best_cnt1), g(r.static_cnt), g(r.best_hyp2), g(r.best_cnt2)
q = lambda x: np.percentile(x, [0, 10, 50, 90, 100]).tolist()
print("matches", q(m)); print("best_cnt1/m", q(bc1 / np.maximum(m, 1))); print("best_hyp1", q(bh1))
print("static", q(sc)); print("best_cnt2/static", q(bc2 / np.maximum(sc, 1))); print("best_hyp2", q(bh2))
print("level2 winner counts every point:", float((bc2 == sc).mean()), " best_hyp2 < 16:", float((bh2 < 16).mean()), " < 64:", float((bh2 < 64).mean()))
print("status", np.bincount(g(r.status)))
