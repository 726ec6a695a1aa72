"""One rank's share of the multi-GPU symmetric search, timed on ONE GPU (diagnostic): kernel time, executed tiles and
log records for part p of P.  usage: python scripts/part_probe.py P [p ...]   (SLIC_SYM_DEBUG=1 prints the record counts)"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_similarity_search_b200 import _lib, synth
from video_similarity_search_b200.backend import CudaBackend

be = CudaBackend()
lib = _lib.load()
x = be.to_device(synth.config("C3"))
parts = int(sys.argv[1])
which = [int(v) for v in sys.argv[2:]] or [0, parts // 2, parts - 1]
two_phase = os.environ.get("PROBE_TWO_PHASE", "1") == "1" and parts > 1
bests = None
if two_phase:   # what the all-reduce MAX delivers: every part's row bests merged
    unit, ub = be.normalize_rows(x)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for part in range(parts):
        b = torch.empty(x.shape[0], dtype=torch.int32, device=x.device)
        e0.record()
        _lib.call("slic_sym_row_bests", unit.data_ptr(), ub.data_ptr(), x.shape[0], x.shape[1], ub.shape[1], part, parts,
                  b.data_ptr(), None)
        e1.record()
        torch.cuda.synchronize()
        bests = b if bests is None else torch.maximum(bests, b)
    print("phase 1 (row bests of one part, whole call): %.3f ms" % e0.elapsed_time(e1))
lib.slic_profile_screen(1)
for part in which:
    for rep in range(2):
        keys, _ = be.first_neighbors_part(x, part, parts, reduce_max=(lambda t: t.copy_(bests)) if two_phase else None)
        ms, fl, ex = ctypes.c_float(0), ctypes.c_double(0), ctypes.c_double(0)
        lib.slic_last_screen_time(ctypes.byref(ms), ctypes.byref(fl))
        lib.slic_last_screen_exec_flop(ctypes.byref(ex))
    tiles = ex.value / (2.0 * 256 * 256 * 512)
    st = be.last_stats.cpu().tolist()
    print("part %d/%d: kernel %.3f ms, %.0f tiles, %.1f ns/tile, re-ranked %d" % (part, parts, ms.value, tiles, ms.value * 1e6 / tiles, st[0]), flush=True)
