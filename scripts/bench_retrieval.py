"""Retrieval top-k timing (diagnostic): tensor-core screen path vs the exact kernels.
usage: python scripts/bench_retrieval.py [C2|C5|QxNxDxk]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_similarity_search_b200 import _lib
from video_similarity_search_b200.backend import CudaBackend

be = CudaBackend()
lib = _lib.load()
which = sys.argv[1] if len(sys.argv) > 1 else "C2"
shapes = {"C2": (3783, 9537, 512, 50, 101), "C5": (100000, 1000000, 1024, 50, 1000)}
if which in shapes:
    nq, n, d, k, kc = shapes[which]
else:
    nq, n, d, k = [int(v) for v in which.split("x")]
    kc = max(2, n // 1000)
g = torch.Generator(device="cuda").manual_seed(0)
centres = torch.randn(kc, d, device="cuda", generator=g)
def mixture(m):
    out = torch.empty(m, d, device="cuda")
    for s in range(0, m, 65536):
        e = min(m, s + 65536)
        lab = torch.randint(0, kc, (e - s,), device="cuda", generator=g)
        out[s:e] = centres[lab] + torch.randn(e - s, d, device="cuda", generator=g)
    return out
x, q = mixture(n), mixture(nq)
ux, xb = be.normalize_rows(x)
uq, qb = be.normalize_rows(q)
torch.cuda.synchronize()

def timeit(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out

lib.slic_profile_screen(1)
ms_tc, (ti, tv) = timeit(lambda: be.topk_cosine(uq, ux, k, q_f16=qb, x_f16=xb), 3)
import ctypes
ms = ctypes.c_float(0); fl = ctypes.c_double(0)
lib.slic_last_screen_time(ctypes.byref(ms), ctypes.byref(fl))
lib.slic_profile_screen(0)
stats = be.last_stats.cpu().tolist()
print("%s: Q=%d N=%d D=%d k=%d" % (which, nq, n, d, k))
print("  tensor-core path: %.3f ms total; screen kernel %.3f ms = %.1f TFLOP/s; candidates re-ranked/row %.1f, rows to exact finisher %d, compactions %d"
      % (ms_tc, ms.value, fl.value / ms.value / 1e9, stats[0] / nq, stats[1], stats[2]))
if nq * n <= 4e9:
    ms_ex, (ei, ev) = timeit(lambda: be.topk_cosine(uq, ux, k), 2)
    print("  exact kernels:    %.3f ms; identical indices: %s" % (ms_ex, bool((ei == ti).all())))
