"""Host-to-device copy rate of pinned host memory allocated with cudaHostAllocDefault / Portable / WriteCombined
(bench.py's e2e leg is bound by exactly this copy)."""
import ctypes as C, time, sys
import numpy as np, torch
rt = C.CDLL("libcudart.so.12")
def host_alloc(nbytes, flags):
    p = C.c_void_p()
    rc = rt.cudaHostAlloc(C.byref(p), C.c_size_t(nbytes), C.c_uint(flags))
    assert rc == 0, rc
    buf = (C.c_uint8 * nbytes).from_address(p.value)
    return p, torch.from_numpy(np.frombuffer(buf, dtype=np.uint8))
torch.cuda.init()
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 31
dev = torch.empty(N, dtype=torch.uint8, device="cuda")
for name, flags in (("default", 0), ("portable", 1), ("write-combined", 4), ("torch pin_memory", None)):
    if flags is None:
        h = torch.empty(N, dtype=torch.uint8).pin_memory()
    else:
        p, h = host_alloc(N, flags)
    t0 = time.perf_counter(); h[::4096] = 1; t_touch = time.perf_counter() - t0
    best = 0
    for rep in range(5):
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); dev.copy_(h, non_blocking=True); e1.record(); e1.synchronize()
        best = max(best, N / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    print(f"{name:18s} is_pinned={h.is_pinned()}  H2D {best:6.1f} GB/s   (touch {t_touch*1e3:.0f} ms)", flush=True)
    if flags is not None:
        rt.cudaFreeHost(p)
