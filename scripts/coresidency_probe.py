"""Does a small kernel on another stream become resident while the persistent screen kernel runs?
Thread 1 runs the level-0 self-search of C3 (~22 ms, 148 persistent CTAs); thread 2, 5 ms later, launches the
normalise kernel (and a 1-element torch fill) on its own stream and reports when they finished relative to the
start of the search.  Finished long before the search ends -> co-resident; at its end -> serialised behind it."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_similarity_search_b200 import synth
from video_similarity_search_b200.backend import CudaBackend

be = CudaBackend()
x = be.to_device(synth.config("C3"))
small = torch.randn(4096, 512, device=be.device)
one = torch.zeros(1, device=be.device)
be.first_neighbors(x)
be.normalize_rows(small)
torch.cuda.synchronize()
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
for trial in range(3):
    res = {}
    t0 = time.perf_counter()

    def search():
        with torch.cuda.stream(sa):
            be.first_neighbors(x)
            sa.synchronize()
        res["search"] = time.perf_counter() - t0

    def guest():
        time.sleep(0.008)
        with torch.cuda.stream(sb):
            res["guest_start"] = time.perf_counter() - t0
            one.fill_(1.0)
            sb.synchronize()
            res["fill"] = time.perf_counter() - t0
            be.normalize_rows(small)
            sb.synchronize()
            res["normalize"] = time.perf_counter() - t0

    th = [threading.Thread(target=search), threading.Thread(target=guest)]
    [t.start() for t in th]
    [t.join() for t in th]
    print("trial", trial, {k: round(v * 1e3, 2) for k, v in sorted(res.items(), key=lambda kv: kv[1])}, flush=True)
