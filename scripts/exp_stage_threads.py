"""Host threads that stage a pageable source into the pinned upload buffers (SLIC_STAGE_THREADS, read per call) against
the end-to-end time of FINCH(plain numpy array) at C3 (diagnostic).  usage: python scripts/exp_stage_threads.py [C3]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from video_similarity_search_b200 import synth
from video_similarity_search_b200.backend import CudaBackend
from video_similarity_search_b200.clustering.finch import FINCH

be = CudaBackend()
x = synth.config(sys.argv[1] if len(sys.argv) > 1 else "C3")
xp = torch.from_numpy(x).pin_memory().numpy()
print("host cores:", os.cpu_count(), flush=True)
for _ in range(3):
    FINCH(x, backend=be, verbose=False)
for threads in (8, 4, 12, 16, 24, 8):
    os.environ["SLIC_STAGE_THREADS"] = str(threads)
    t = []
    for _ in range(8):
        t0 = time.perf_counter(); FINCH(x, backend=be, verbose=False); t.append((time.perf_counter() - t0) * 1e3)
    print("%2d staging threads: pageable source %.2f ms median (min %.2f)" % (threads, float(np.median(t)), min(t)), flush=True)
t = []
for _ in range(8):
    t0 = time.perf_counter(); FINCH(xp, backend=be, verbose=False); t.append((time.perf_counter() - t0) * 1e3)
print("pinned source: %.2f ms median (min %.2f)" % (float(np.median(t)), min(t)), flush=True)
for nthreads in (1, 8, 16):
    dst = np.empty_like(x)
    import threading
    def work(i, k):
        n = len(x); a, b = n * i // k, n * (i + 1) // k
        np.copyto(dst[a:b], x[a:b])
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(i, nthreads)) for i in range(nthreads)]
    [w.start() for w in th]; [w.join() for w in th]
    dt = time.perf_counter() - t0
    print("numpy copy of the 492 MB matrix with %2d threads: %.1f ms = %.1f GB/s" % (nthreads, dt * 1e3, x.nbytes / dt / 1e9), flush=True)
