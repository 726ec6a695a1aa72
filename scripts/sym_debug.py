"""Symmetric screen on small inputs: records logged, fullest region (diagnostic)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from video_similarity_search_b200 import _lib, synth
from video_similarity_search_b200.backend import CudaBackend
be = CudaBackend(); lib = _lib.load()
for n, d, k, seed in ((16384, 64, 40, 1), (20011, 96, 0, 2), (50000, 512, 300, 3), (24000, 640, 50, 5)):
    x = synth.gaussian_mixture(n, d, k, seed) if k else np.random.default_rng(seed).standard_normal((n, d)).astype(np.float32)
    unit, ub = be.normalize_rows(be.to_device(x))
    lib.slic_profile_screen(1)
    idx, dist = be.nn_top1(unit, ub, unit, ub, self_offset=0)
    ms, fl, ex = ctypes.c_float(0), ctypes.c_double(0), ctypes.c_double(0)
    lib.slic_last_screen_time(ctypes.byref(ms), ctypes.byref(fl)); lib.slic_last_screen_exec_flop(ctypes.byref(ex))
    st = be.last_stats.cpu().tolist()
    print(n, d, k, "exec/alg %.2f" % (ex.value / fl.value), "stats", st, "logged/row %.1f" % (st[3] * 16.0 / n))
