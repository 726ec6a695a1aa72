"""Pipeline trace of the level-0 screen (slic_screen_trace): who waits for whom (diagnostic).
usage: [SLIC_SCREEN_SYM=0] python scripts/screen_trace.py [C3|NxD]"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_similarity_search_b200 import _lib, synth
from video_similarity_search_b200.backend import CudaBackend

be = CudaBackend(); lib = _lib.load()
which = sys.argv[1] if len(sys.argv) > 1 else "C3"
if which.endswith("L1"):   # level 1 of a config: the float64 centroids of its level-0 clusters
    x0 = be.to_device(synth.config(which[:-2]))
    nn0, _, _ = be.first_neighbors(x0)
    lab0, c0 = be.components(nn0)
    x = be.cluster_sums(x0, lab0, c0)[2]
elif which in synth.CONFIGS:
    x = be.to_device(synth.config(which))
else:
    n, d = [int(v) for v in which.split("x")]
    g = torch.Generator(device="cuda").manual_seed(0)
    cen = torch.randn(max(2, n // 600), d, device="cuda", generator=g)
    x = cen[torch.randint(0, cen.shape[0], (n,), device="cuda", generator=g)] + torch.randn(n, d, device="cuda", generator=g)
unit, ub = be.normalize_rows(x)
be.nn_top1(unit, ub, unit, ub, self_offset=0)
torch.cuda.synchronize()
lib.slic_screen_trace(1, None)
lib.slic_profile_screen(1)
be.nn_top1(unit, ub, unit, ub, self_offset=0)
torch.cuda.synchronize()
c = (ctypes.c_uint64 * 12)()
lib.slic_screen_trace(0, ctypes.addressof(c))
ms, fl = ctypes.c_float(0), ctypes.c_double(0)
lib.slic_last_screen_time(ctypes.byref(ms), ctypes.byref(fl))
ctas, pairs = 148, 74
print("%s screen %.3f ms" % (which, ms.value))
print("  TMA producer waited for a free stage : %5.1f %% of the MMA issuer's time" % (100.0 * c[0] / ctas / (c[3] / pairs)))
print("  MMA issuer waited for an accumulator : %5.1f %%  (epilogue-bound)" % (100.0 * c[1] / c[3]))
print("  MMA issuer waited for operands       : %5.1f %%  (TMA / L2-bound)" % (100.0 * c[2] / c[3]))
print("  epilogue warp waited for the MMAs    : %5.1f %% of its time" % (100.0 * c[4] / max(c[5], 1)))
print("  column-role 32x32 chunks that took the slow path (epilogue warp 0 of every CTA): %d of %d = %.2f %%" % (c[6], c[7], 100.0 * c[6] / max(c[7], 1)))
print("  MMA issuer waited for a unit id       : %5.1f %%  (scheduler / pre-pass hand-over); units %d" % (100.0 * c[8] / c[3], c[10]))
print("  epilogue warp 0 waited for a unit id  : %5.1f %% of its time" % (100.0 * c[9] / max(c[5], 1)))
print("  raw", list(c))
