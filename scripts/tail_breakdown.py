"""Serial tail of a FINCH step (everything after the level-0 search): per backend call, wall clock with a device
sync after each, through the Python level loop (the native driver makes the same calls).  Diagnostic only."""
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
        self.be, self.t, self.log = be, {}, []

    def __getattr__(self, name):
        if name in ("finch_native", "finch_host"):
            raise AttributeError(name)
        f = getattr(self.be, name)
        if not callable(f):
            return f

        def wrap(*a, **k):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            r = f(*a, **k)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            self.t[name] = self.t.get(name, 0) + dt
            shape = tuple(a[0].shape) if a and hasattr(a[0], "shape") else ()
            self.log.append((name, shape, round(dt * 1e3, 3)))
            return r
        return wrap


for it in range(3):
    tb = Timed(be)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c, num, _ = fm.FINCH(x, backend=tb, verbose=False)
    torch.cuda.synchronize(); tot = time.perf_counter() - t0
    print("iter", it, "total %.2f ms" % (tot * 1e3), {k: round(v * 1e3, 2) for k, v in tb.t.items()}, num)
print(tb.log)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c, num, _ = fm.FINCH(x, backend=be, verbose=False)
    torch.cuda.synchronize(); tot = time.perf_counter() - t0
    torch.cuda.synchronize(); t0 = time.perf_counter()
    be.first_neighbors(x)
    torch.cuda.synchronize(); nn = time.perf_counter() - t0
    print("native driver: total %.2f ms, level-0 search alone %.2f ms" % (tot * 1e3, nn * 1e3))
