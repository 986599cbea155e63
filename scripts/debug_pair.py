"""Bring-up check of the CTA-pair match kernel (variant 7) against the single-CTA V-space kernel (variant 0)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import evenvizion_b200 as evz
from evenvizion_b200 import synth
P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
eng = evz.GeometryEngine(0)
ch = synth.make_chain(P + 1, N, seed=3, device="cuda")
st = eng.ingest(ch["desc"], ch["coords"])
outs = {}
for v in (0, 7):
    eng.set_option(2, v)
    outs[v] = eng.match(st, list(range(1, P + 1)), list(range(0, P)))
    torch.cuda.synchronize()
    print("variant", v, "done", flush=True)
q0 = int(st.row_off_h[1])
a, b = outs[0], outs[7]
print("idx equal:", torch.equal(a.top2_idx[q0:], b.top2_idx[q0:]), " d2 equal:", torch.equal(a.top2_d2[q0:], b.top2_d2[q0:]))
if not torch.equal(a.top2_idx[q0:], b.top2_idx[q0:]):
    bad = (a.top2_idx[q0:] != b.top2_idx[q0:]).any(1).nonzero().flatten()
    print("rows differing:", bad.numel(), bad[:20].tolist())
    for r in bad[:8].tolist():
        print(r, a.top2_idx[q0 + r].tolist(), b.top2_idx[q0 + r].tolist(), a.top2_d2[q0 + r].tolist(), b.top2_d2[q0 + r].tolist())
