"""Rank-0-driven multi-GPU FINCH (slic_finch_multi: one process, one worker thread per device) on the bench workload:
wall time per call from a pinned and from a pageable host matrix, the library's own CUDA-event timeline of the last call,
and the partition against the single-GPU host entry (diagnostic; python scripts/multi_single_process.py [C3] [devices])."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from video_similarity_search_b200 import synth
from video_similarity_search_b200.backend import CudaBackend
from video_similarity_search_b200.clustering.finch import FINCH

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
ndev = int(sys.argv[2]) if len(sys.argv) > 2 else torch.cuda.device_count()
x = synth.config(name)
xp = torch.from_numpy(x).pin_memory().numpy()
be1, bem = CudaBackend(), CudaBackend()
c1, num1, _ = FINCH(x, backend=be1, verbose=False)
bem.enable_multi_gpu(devices=list(range(ndev)), max_rows=max(len(x), 1 << 18))
for label, src in (("pinned", xp), ("pageable", x)):
    for _ in range(3):
        cm, numm, _ = FINCH(src, backend=bem, verbose=False)
    assert numm == num1 and np.array_equal(cm, c1)
    t = []
    for _ in range(5):
        t0 = time.perf_counter(); FINCH(src, backend=bem, verbose=False); t.append((time.perf_counter() - t0) * 1e3)
    up, search, total = bem.multi_gpu_timeline()
    print("%d GPUs, %s source: wall %.2f ms (min %.2f) | device 0 timeline: upload+forward %.2f, normalise+search %.2f, whole %.2f"
          % (ndev, label, float(np.median(t)), min(t), up, search, total), flush=True)
for label, src in (("pinned", xp), ("pageable", x)):
    t = []
    for _ in range(5):
        t0 = time.perf_counter(); FINCH(src, backend=be1, verbose=False); t.append((time.perf_counter() - t0) * 1e3)
    print("1 GPU, %s source: wall %.2f ms (min %.2f)" % (label, float(np.median(t)), min(t)), flush=True)
bem.disable_multi_gpu()
