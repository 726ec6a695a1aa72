"""Per-call wall time of the resident FINCH step while the caller keeps the previous result alive (diagnostic for the
page-locked result pool of backend.py).  usage: python scripts/diag_labels.py [C3]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_similarity_search_b200 import synth
from video_similarity_search_b200.backend import CudaBackend
from video_similarity_search_b200.clustering.finch import FINCH

be = CudaBackend()
x = be.to_device(synth.config(sys.argv[1] if len(sys.argv) > 1 else "C3"))
for hold in (False, True, True):
    out = None
    for i in range(8):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = FINCH(x, backend=be, verbose=False)
        torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 1e3
        if hold:
            out = r
        del r
        print("hold=%s call %d: %.2f ms, pool buffers %d (%s MB)" % (hold, i, ms, len(be._results._entries),
              ",".join("%.0f" % (e[0].numel() / 1e6) for e in be._results._entries)), flush=True)
t0 = time.perf_counter(); t = torch.empty(32 << 20, dtype=torch.uint8, pin_memory=True); print("pinned 32 MB alloc: %.2f ms" % ((time.perf_counter() - t0) * 1e3))
t0 = time.perf_counter(); t2 = torch.empty(8 << 20, dtype=torch.uint8, pin_memory=True); print("pinned 8 MB alloc: %.2f ms" % ((time.perf_counter() - t0) * 1e3))
