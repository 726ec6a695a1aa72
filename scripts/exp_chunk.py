"""Column-chunk width of the symmetric screen's chunk-major unit order (SLIC_SYM_CHUNK_TILES, in 256-column blocks) against
screen time at C3 - the A rows of a unit are re-read from HBM once per column chunk, so DRAM traffic falls as the chunk
grows while the chunk's B tiles (16 MB per 64 blocks) must stay shared through L2 (diagnostic).
usage: python scripts/exp_chunk.py [C3] [chunk,unit ...]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_similarity_search_b200 import _lib, synth
from video_similarity_search_b200.backend import CudaBackend

be = CudaBackend(); lib = _lib.load()
x = be.to_device(synth.config(sys.argv[1] if len(sys.argv) > 1 else "C3"))
unit, ub = be.normalize_rows(x)
variants = [tuple(int(v) for v in a.split(",")) for a in sys.argv[2:]] or [(64, 64), (96, 48), (128, 64), (192, 64), (256, 64), (64, 64)]
ref = None
lib.slic_profile_screen(1)
for chunk, ulen in variants:
    os.environ["SLIC_SYM_CHUNK_TILES"] = str(chunk)
    os.environ["SLIC_SYM_UNIT_TILES"] = str(ulen)
    times = []
    for it in range(6):
        idx, dist = be.nn_top1(unit, ub, unit, ub, self_offset=0)
        torch.cuda.synchronize()
        ms, fl = ctypes.c_float(0), ctypes.c_double(0)
        lib.slic_last_screen_time(ctypes.byref(ms), ctypes.byref(fl))
        if it >= 2:
            times.append(ms.value)
    st = be.last_stats.cpu().tolist()
    if ref is None:
        ref = idx.clone()
    print("chunk %3d blocks, units %2d: screen %.3f ms (min %.3f) | logged/row %.1f re-ranked/row %.2f | equal to first variant: %s"
          % (chunk, ulen, sum(times) / len(times), min(times), st[3] * 16.0 / x.shape[0], st[0] / x.shape[0], torch.equal(idx, ref)),
          flush=True)
