"""Short driver for ncu: a few FINCH steps on the bench workload (no timing claims are taken from this)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_similarity_search_b200 import synth
from video_similarity_search_b200.backend import CudaBackend
from video_similarity_search_b200.clustering.finch import FINCH

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
name = sys.argv[2] if len(sys.argv) > 2 else "C3"
be = CudaBackend()
x = be.to_device(synth.config(name))
torch.cuda.synchronize()
for _ in range(steps):
    c, num, _ = FINCH(x, backend=be, verbose=False)
torch.cuda.synchronize()
print(num)
