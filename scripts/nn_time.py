"""Level-0 first-neighbour stage alone: screen kernel time, algorithmic and executed flop (diagnostic).
usage: python scripts/nn_time.py [C3|C5|NxD]"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_similarity_search_b200 import _lib, synth
from video_similarity_search_b200.backend import CudaBackend

be = CudaBackend(); lib = _lib.load()
which = sys.argv[1] if len(sys.argv) > 1 else "C3"
if which in synth.CONFIGS:
    x = be.to_device(synth.config(which))
else:
    n, d = [int(v) for v in which.split("x")]
    g = torch.Generator(device="cuda").manual_seed(0)
    cen = torch.randn(max(2, n // 600), d, device="cuda", generator=g)
    x = cen[torch.randint(0, cen.shape[0], (n,), device="cuda", generator=g)] + torch.randn(n, d, device="cuda", generator=g)
unit, ub = be.normalize_rows(x)
torch.cuda.synchronize()
lib.slic_profile_screen(1)
for it in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    idx, dist = be.nn_top1(unit, ub, unit, ub, self_offset=0)
    torch.cuda.synchronize(); wall = (time.perf_counter() - t0) * 1e3
    ms, fl, ex = ctypes.c_float(0), ctypes.c_double(0), ctypes.c_double(0)
    lib.slic_last_screen_time(ctypes.byref(ms), ctypes.byref(fl)); lib.slic_last_screen_exec_flop(ctypes.byref(ex))
    st = be.last_stats.cpu().tolist()
    print("%s: nn_top1 wall %.3f ms | screen %.3f ms: algorithmic %.1f TFLOP/s, executed %.1f TFLOP/s (%.1f%% of the square) | "
          "re-ranked/row %.2f, logged/row %.1f, exact-finished rows %d" % (which, wall, ms.value, fl.value / ms.value / 1e9, ex.value / ms.value / 1e9,
                                                          100 * ex.value / fl.value, st[0] / x.shape[0], st[3] * 16.0 / x.shape[0], st[1]))
