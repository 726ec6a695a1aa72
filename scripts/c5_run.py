"""BASELINE configs[4] on ONE GPU (scale sweep N = 1 000 000 x D = 1 024, d_pad > 512: the streaming pair kernel):
FINCH full hierarchy timed with CUDA events, first neighbours checked against the oracle on sampled rows, and the
whole partition against the oracle's levels >= 1 run from the GPU's level-0 neighbours.  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from video_similarity_search_b200 import synth
from video_similarity_search_b200.backend import CudaBackend
from video_similarity_search_b200.clustering.finch import FINCH

sample = int(sys.argv[1]) if len(sys.argv) > 1 else 512
be = CudaBackend()
t0 = time.perf_counter()
x = synth.config("C5")
gen_s = time.perf_counter() - t0
xd = be.to_device(x)
torch.cuda.synchronize()
FINCH(xd, backend=be, verbose=False)            # warm-up (memory pool, module load)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
c, num, _ = FINCH(xd, backend=be, verbose=False)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
e0.record()
nn, dist, _ = be.first_neighbors(xd)
e1.record()
torch.cuda.synchronize()
ms_nn = e0.elapsed_time(e1)
out = {"workload": "C5 FINCH N=1000000 x D=1024 (K=1000, seed 0), 1 GPU", "finch_ms": ms, "nn_stage_ms": ms_nn,
       "partitions": [int(v) for v in num], "nn_algorithmic_tflops": 2.0 * 1e6 * 1e6 * 1024 / (ms_nn * 1e-3) / 1e12,
       "generate_s": gen_s}
if sample > 0:
    from oracle import finch_oracle as fo
    rows = np.linspace(0, len(x) - 1, sample).astype(np.int64)
    t0 = time.perf_counter()
    enn, _, gap = fo.first_neighbors_blocked(x, rows=rows)
    out["oracle_nn_sample_s"] = time.perf_counter() - t0
    nn_h = nn.cpu().numpy().astype(np.int64)
    clear = gap > 2e-6
    out["first_neighbors_equal_on_sample"] = bool(np.array_equal(nn_h[rows][clear], enn[clear]))
    out["sample_rows"] = int(sample)
    out["tie_rows_in_sample"] = int((~clear).sum())
    t0 = time.perf_counter()
    co, no, _ = fo.finch(x, initial_rank=nn_h)
    out["oracle_levels_s"] = time.perf_counter() - t0
    out["partition_equals_oracle"] = bool(no == num and np.array_equal(co, c))
print(json.dumps(out))
