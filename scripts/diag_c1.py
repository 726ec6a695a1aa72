"""Diagnostic: where does the CUDA FINCH differ from the reference golden at C1, and is it explained by tie rows?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import finch_oracle as fo
from video_similarity_search_b200 import synth
from video_similarity_search_b200.backend import CudaBackend
from video_similarity_search_b200.clustering.finch import FINCH

be = CudaBackend()
g = np.load("tests/golden/c1_9537x512.npz")
x = synth.config("C1")
nn, d, unit = be.first_neighbors(be.to_device(x))
nn = nn.cpu().numpy()
enn, ed, gap = fo.first_neighbors_blocked(x)
diff = np.nonzero(nn != g["nn_level0"])[0]
print("rows where nn differs from reference golden:", diff.tolist())
for r in diff:
    print(" row", r, "gpu", nn[r], "ref", g["nn_level0"][r], "gap", gap[r], "oracle_blocked", enn[r])
print("rows with gap < 2e-6:", np.nonzero(gap < 2e-6)[0].tolist(), gap[gap < 2e-6])
ms = be.min_sim(be.to_device(nn.astype(np.int32)), unit, d)
print("min_sim gpu", repr(ms), "ref", repr(g["min_sim"]))
c, num, _ = FINCH(x, backend=be, verbose=False)
print("num", num, g["num_clust"].tolist())
for lvl in range(c.shape[1]):
    bad = np.nonzero(c[:, lvl] != g["c"][:, lvl])[0]
    print("level", lvl, "label mismatches", len(bad), bad[:10].tolist())
