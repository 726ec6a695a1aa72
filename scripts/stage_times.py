"""Where does a FINCH step spend its time?  Wall-clock per stage with a device sync after each (diagnostic only)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_similarity_search_b200 import synth
from video_similarity_search_b200.backend import CudaBackend
from video_similarity_search_b200.clustering import finch as fm

be = CudaBackend()
x = be.to_device(synth.config(sys.argv[1] if len(sys.argv) > 1 else "C3"))
torch.cuda.synchronize()

class Timed:
    def __init__(self, be):
        self.be, self.t = be, {}
    def __getattr__(self, name):
        f = getattr(self.be, name)
        if not callable(f):
            return f
        def wrap(*a, **k):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            r = f(*a, **k)
            torch.cuda.synchronize(); self.t[name] = self.t.get(name, 0) + time.perf_counter() - t0
            return r
        return wrap

for it in range(3):
    tb = Timed(be)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c, num, _ = fm.FINCH(x, backend=tb, verbose=False)
    torch.cuda.synchronize(); tot = time.perf_counter() - t0
    print("iter", it, "total %.2f ms" % (tot * 1e3), {k: round(v * 1e3, 2) for k, v in tb.t.items()}, num)
