"""How long is everything AFTER the level-0 search?  FINCH with the level-0 neighbours cached (diagnostic)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_similarity_search_b200 import synth
from video_similarity_search_b200.backend import CudaBackend
from video_similarity_search_b200.clustering.finch import FINCH

be = CudaBackend()
x = be.to_device(synth.config(sys.argv[1] if len(sys.argv) > 1 else "C3"))
cached = be.first_neighbors(x)
torch.cuda.synchronize()
for it in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c, num, _ = FINCH(x, backend=be, verbose=False, first_neighbors=lambda m: cached)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print("tail (levels >= 1, components, means, D2H): %.3f ms" % ((t1 - t0) * 1e3), num)
