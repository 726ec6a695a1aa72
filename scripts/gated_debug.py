"""Host-entry (gated upload) smoke: slic_finch_host on a synthetic matrix; prints the partition or the error text
(which carries the post-mortem of a timed-out wait).  Variants via SLIC_GATED_CHUNKS / SLIC_SCREEN_SYM."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from video_similarity_search_b200 import synth
from video_similarity_search_b200.backend import CudaBackend

n, d = int(sys.argv[1]), int(sys.argv[2])
be = CudaBackend()
x = synth.gaussian_mixture(n, d, 120, 17)
dev = be.to_device(x)
c_dev, num_dev, _ = be.finch_native(dev)
torch.cuda.synchronize()
print("resident:", num_dev, flush=True)
t0 = time.perf_counter()
try:
    c, num, _ = be.finch_host(x)
    print("host entry: %.1f ms" % ((time.perf_counter() - t0) * 1e3), num, "equal:", np.array_equal(c, c_dev.cpu().numpy()), flush=True)
except Exception as e:
    print("host entry FAILED after %.1f s: %s" % (time.perf_counter() - t0, e), flush=True)
    os._exit(0)
