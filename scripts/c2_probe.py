"""Per-call times of the configs[1] top-50 retrieval in the situations bench.py creates (diagnostic)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from video_similarity_search_b200 import synth
from video_similarity_search_b200.backend import CudaBackend
from video_similarity_search_b200.clustering.finch import FINCH
be = CudaBackend()
tr, _, te, _ = synth.c2_retrieval()
def per_call(label, fn, reps=6):
    rows = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        rows.append("%.2f/%.2f" % (e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3))
    print("%-40s %s" % (label, " ".join(rows)), flush=True)
q, x = be.to_device(te), be.to_device(tr)
per_call("top50 fresh process", lambda: be.topk_neighbors(q, x, 50))
xh = synth.config("C3")
xd = be.to_device(xh)
per_call("FINCH resident C3", lambda: FINCH(xd, backend=be, verbose=False), 3)
per_call("top50 after FINCH", lambda: be.topk_neighbors(q, x, 50))
per_call("FINCH pageable C3", lambda: FINCH(xh, backend=be, verbose=False), 3)
per_call("top50 after pageable FINCH", lambda: be.topk_neighbors(q, x, 50))
q2, x2 = be.to_device(te), be.to_device(tr)
per_call("top50 on fresh copies", lambda: be.topk_neighbors(q2, x2, 50))
