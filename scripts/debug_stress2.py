import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import evenvizion_b200 as evz
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests")); import test_gpu_match as T
eng = evz.GeometryEngine(0)
SYNC = int(sys.argv[1]) if len(sys.argv) > 1 else 1
orig = eng.match
n = [0]
def match(st, pq, pt, **kw):
    n[0] += 1
    print("call", n[0], "variant", eng._v, "pairs", len(pq), flush=True)
    r = orig(st, pq, pt, **kw)
    if SYNC:
        torch.cuda.synchronize(); print("   done", flush=True)
    return r
so = eng.set_option
def set_option(o, v):
    eng._v = v; so(o, v)
eng._v = 0
eng.match = match; eng.set_option = set_option
T.test_match_ragged_stress_against_device_reference(eng)
print("TEST OK")
