"""Pipeline counters (slic_screen_trace) of the level-0 screen inside slic_finch_host - the GATED launch that starts before
the embeddings have arrived - next to the same screen on resident data: how much of the e2e overhead is the MMA issuer
waiting for units whose rows are still crossing PCIe, and how much is a slower kernel (diagnostic).
usage: [SLIC_GATED_CHUNKS=c] python scripts/gated_trace.py [C3]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_similarity_search_b200 import _lib, synth
from video_similarity_search_b200.backend import CudaBackend

be = CudaBackend(); lib = _lib.load()
x = synth.config(sys.argv[1] if len(sys.argv) > 1 else "C3")
xp = torch.from_numpy(x).pin_memory().numpy()
dev = be.to_device(x)


def report(label, fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    lib.slic_screen_trace(1, None)
    lib.slic_profile_screen(1)
    fn()
    torch.cuda.synchronize()
    c = (ctypes.c_uint64 * 12)()
    lib.slic_screen_trace(0, ctypes.addressof(c))
    lib.slic_profile_screen(0)
    pairs = 74
    cyc = c[3] / pairs
    print("%-28s MMA issuer %.0f kcycles per pair | waits: accumulator %.1f %%, operands %.1f %%, unit id (gates / pre-pass) %.1f %% "
          "= %.0f kcycles | issuing %.0f kcycles | units %d | raw %s"
          % (label, cyc / 1e3, 100.0 * c[1] / c[3], 100.0 * c[2] / c[3], 100.0 * c[8] / c[3], c[8] / pairs / 1e3,
             (c[3] - c[1] - c[2] - c[8]) / pairs / 1e3, c[10], list(c)), flush=True)


report("resident (both levels)", lambda: be.finch_native(dev))
report("host matrix, gated upload", lambda: be.finch_host(xp))
report("host matrix, pageable", lambda: be.finch_host(x))
