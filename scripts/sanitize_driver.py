"""Small-shape pass over every kernel family for compute-sanitizer (memcheck / racecheck / synccheck / initcheck).
Shapes are chosen so that each code path runs at least once: symmetric tcgen05 screen (16 384 x 64), full-square pair
kernel (rectangular top-1), top-k screen, exact kernels, components with and without the min_sim cut, means, the
device-side small-levels kernel, masks, grouping, scatter, metrics, centring.  Results are checked against the oracle
so that a "clean" report is about a run that computed the right thing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import finch_oracle as fo
from oracle import masks_oracle as mo
from video_similarity_search_b200 import cluster_io, metrics, synth
from video_similarity_search_b200.backend import CudaBackend
from video_similarity_search_b200.clustering import cluster_masks as cm
from video_similarity_search_b200.clustering.finch import FINCH

be = CudaBackend()
# K1 symmetric screen + device-side hierarchy (device matrix and host matrix entries)
x = synth.gaussian_mixture(16384, 64, 40, 1)
c, num, _ = FINCH(be.to_device(x), backend=be, verbose=False)
co, no, _ = fo.finch(x)
assert num == no and np.array_equal(c, co), (num, no)
ch, numh, _ = FINCH(x, backend=be, verbose=False)
assert numh == no and np.array_equal(ch, co)
# K1 rectangular top-1 (full-square pair kernel) and exact kernel
xd = be.to_device(x)
unit, ub = be.normalize_rows(xd)
q = slice(100, 2148)
i_tc, d_tc = be.nn_top1(unit[q], ub[q], unit, ub, self_offset=100)
i_ex, d_ex = be.nn_exact_top1(unit[q], unit, self_offset=100)
assert torch.equal(i_tc, i_ex)
# K1 top-k screen vs exact
tr, ytr, te, yte = synth.c2_retrieval()
te, tr = te[:1024, :128].copy(), tr[:, :128].copy()
i1, v1 = be.topk_neighbors(be.to_device(te), be.to_device(tr), 10)
uq, _ = be.normalize_rows(be.to_device(te), want_f16=False)
ux, _ = be.normalize_rows(be.to_device(tr), want_f16=False)
i2, v2 = be.topk_cosine(uq, ux, 10)
assert torch.equal(i1, i2)
hits = be.hit_at_k(i1, be.to_device(yte[:1024]), be.to_device(ytr), [1, 5, 10])
# small FINCH with the min_sim filter (exact kernel, components with sibling pairs, small-levels kernel)
xs = synth.gaussian_mixture(3000, 128, 30, 7)
cs, nums, _ = FINCH(xs, backend=be, verbose=False)
cso, nso, _ = fo.finch(xs)
assert nums == nso and np.array_equal(cs, cso)
_, _, req = FINCH(xs, req_clust=25, backend=be, verbose=False)
assert len(np.unique(req)) == 25
# K3 / K4 / metrics / hand-over
lab = cs[:, 0]
mask = cm.queue_positive_mask(lab[:64], lab[64:2112], backend=be).cpu().numpy()
assert np.array_equal(mask, mo.queue_positive_mask(lab[:64], lab[64:2112]))
order, offsets = cm.group_by_label(lab, nums[0], backend=be)
sums, counts, means = be.cluster_sums(be.to_device(xs), be.to_device(lab.astype(np.int32)), nums[0])
np.testing.assert_allclose(means.cpu().numpy(), fo.cluster_means(xs, lab), rtol=0, atol=1e-11)
nmi = metrics.normalized_mutual_info_score(lab % 7, lab, backend=be)
ami = metrics.adjusted_mutual_info_score(lab % 7, lab, backend=be)
out, bad = be.scatter_last_wins(be.to_device(lab.astype(np.int32)), be.to_device(np.arange(3000)[::-1].astype(np.int32).copy()), 3000)
cen = be.center_columns(be.to_device(xs))
dm = be.distance_matrix(uq[:300], ux[:500])
torch.cuda.synchronize()
print("sanitize driver ok", num, nums, hits.tolist(), round(nmi, 4), round(ami, 4))
