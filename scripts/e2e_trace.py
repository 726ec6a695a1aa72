"""Timeline of the host entry point slic_finch_host on the bench workload (diagnostic).
usage: [SLIC_GATED_CHUNKS=c] python scripts/e2e_trace.py [C3]"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_similarity_search_b200 import _lib, synth
from video_similarity_search_b200.backend import CudaBackend

be = CudaBackend()
lib = _lib.load()
x = torch.from_numpy(synth.config(sys.argv[1] if len(sys.argv) > 1 else "C3")).pin_memory()
xn = x.numpy()
dev = x.cuda()
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter(); y = x.to("cuda", non_blocking=True); torch.cuda.synchronize()
    print("H2D alone: %.3f ms" % ((time.perf_counter() - t0) * 1e3))
for _ in range(3):
    t0 = time.perf_counter(); be.finch_native(dev); torch.cuda.synchronize()
    print("resident slic_finch: %.3f ms" % ((time.perf_counter() - t0) * 1e3))
lib.slic_host_trace(1, None)
lib.slic_profile_screen(1)
ms = (ctypes.c_float * 4)()
for _ in range(5):
    t0 = time.perf_counter(); c, num, _ = be.finch_host(xn); wall = (time.perf_counter() - t0) * 1e3
    lib.slic_host_trace(1, ctypes.addressof(ms))
    print("slic_finch_host wall %.3f ms | first copy starts +%.3f, upload lasts %.3f, search done +%.3f, labels back +%.3f" %
          (wall, ms[0], ms[1], ms[2], ms[3]), num)
