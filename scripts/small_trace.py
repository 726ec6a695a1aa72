"""Phase timeline of the device-side level loop (SLIC_SMALL_TRACE=1) + grid-barrier cost probe (diagnostic)."""
import os, sys
os.environ["SLIC_SMALL_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_similarity_search_b200 import synth
from video_similarity_search_b200.backend import CudaBackend
from video_similarity_search_b200.clustering.finch import FINCH
be = CudaBackend()
x = be.to_device(synth.config(sys.argv[1] if len(sys.argv) > 1 else "C3"))
for _ in range(3):
    c, num, _ = FINCH(x, backend=be, verbose=False)
print(num)
