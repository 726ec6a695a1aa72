for c in 8 16 24; do
SLIC_GATED_CHUNKS=$c python - <<PY
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from video_similarity_search_b200 import synth
from video_similarity_search_b200.backend import CudaBackend
from video_similarity_search_b200.clustering.finch import FINCH
be = CudaBackend(); x = synth.config("C3"); xp = torch.from_numpy(x).pin_memory().numpy()
for name, src in (("pageable", x), ("pinned", xp)):
    for _ in range(3): FINCH(src, backend=be, verbose=False)
    t = []
    for _ in range(10):
        t0 = time.perf_counter(); FINCH(src, backend=be, verbose=False); t.append((time.perf_counter() - t0) * 1e3)
    print("chunks $c %s: %.2f ms median (min %.2f)" % (name, float(np.median(t)), min(t)), flush=True)
PY
done
