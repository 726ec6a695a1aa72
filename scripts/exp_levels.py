"""Per-stage timings (CUDA events, diagnostic): level-0 search, level-1 search (float64 centroids of level 0), full FINCH,
and the tail with the level-0 neighbours cached.  Environment knobs of the screen (SLIC_SYM_*) are read once per
process, so variants are separate runs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_similarity_search_b200 import synth
from video_similarity_search_b200.backend import CudaBackend
from video_similarity_search_b200.clustering.finch import FINCH

be = CudaBackend()
wl = sys.argv[1] if len(sys.argv) > 1 else "C3"
x = be.to_device(synth.config(wl))


def timed(fn, reps=5):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


nn, dist, unit = be.first_neighbors(x)
lab, c0 = be.components(nn)
sums, counts, means = be.cluster_sums(x, lab, c0)
print("knobs:", {k: v for k, v in os.environ.items() if k.startswith("SLIC_")})
import ctypes
from video_similarity_search_b200 import _lib
lib = _lib.load()


def screen_ms(fn):
    lib.slic_profile_screen(1)
    out = []
    for _ in range(3):
        fn()
        ms, fl = ctypes.c_float(0), ctypes.c_double(0)
        lib.slic_last_screen_time(ctypes.byref(ms), ctypes.byref(fl))
        out.append(round(ms.value, 3))
    lib.slic_profile_screen(0)
    return out


print("level-0 search  (%d x %d f32): %.3f ms" % (x.shape[0], x.shape[1], timed(lambda: be.first_neighbors(x))))
print("   screen kernel alone:", screen_ms(lambda: be.first_neighbors(x)))
if c0 >= 2048:
    print("level-1 search  (%d x %d f64): %.3f ms" % (c0, x.shape[1], timed(lambda: be.first_neighbors(means), 10)))
    print("   stats (reranked, exact rows, compactions, logged/16):", be.last_stats.tolist())
    print("   screen kernel alone:", screen_ms(lambda: be.first_neighbors(means)))
print("K3 level 0 (cluster_sums whole call): %.3f ms" % timed(lambda: be.cluster_sums(x, lab, c0), 10))
print("full FINCH resident: %.3f ms" % timed(lambda: FINCH(x, backend=be, verbose=False)))
cached = (nn, dist, unit)
print("tail (cached level-0 neighbours, incl. D2H of labels): %.3f ms" % timed(lambda: FINCH(x, backend=be, verbose=False, first_neighbors=lambda m: cached)))
c, num, _ = FINCH(x, backend=be, verbose=False)
print("partitions", num)
