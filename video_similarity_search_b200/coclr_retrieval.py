"""Nearest-neighbour retrieval of coclr_classify.py on the K1 kernels.

  coclr_classify.py:784-810 (test_retrieval tail)   centring, L2 normalisation, dot products, five torch.topk
                                                    passes over a materialised [Q, N] matrix, kNN accuracy
  loss/triplet_loss.py:429-447 (pdist / pdist_v2)   in-batch distance matrices built row by row in Python

Here: one centring pre-pass (slic_center_columns), one tensor-core top-50 search with exact re-rank
(slic_topk_cosine_tc; the [Q, N] similarity matrix is never formed) and one hit-counting launch.
"""
import numpy as np
import torch

from . import backend as _backend

KS = [1, 5, 10, 20, 50]


def _f32_on_device(be, t):
    if isinstance(t, torch.Tensor):
        t = t.detach()
    return be.to_device(t, torch.float32)


def retrieval_topk(test_feature, train_feature, k, backend=None):
    """coclr_classify.py:788-796 without `sim`: centre both sides by their own column means, normalise, and return
    (idx int32 [Q, k], similarity float32 [Q, k]) - the k largest dot products per test row, best first."""
    be = backend or _backend.default_backend()
    q = be.center_columns(_f32_on_device(be, test_feature))        # :788
    x = be.center_columns(_f32_on_device(be, train_feature))       # :789
    idx, dist = be.topk_neighbors(q, x, k)                         # :792-796 (normalise + dot product) and :807 (topk)
    return idx, 1.0 - dist


def nn_retrieval_accuracy(test_feature, test_label, train_feature, train_label, ks=KS, backend=None):
    """coclr_classify.py:784-810: [acc@k for k in ks], acc@k = fraction of test rows whose label occurs among the
    labels of their k most similar train rows.  Prints the reference's '%dNN acc = %.4f' lines."""
    be = backend or _backend.default_backend()
    kmax = min(max(ks), int(train_feature.shape[0]))
    idx, _ = retrieval_topk(test_feature, train_feature, kmax, backend=be)
    ql = be.to_device(torch.as_tensor(test_label).detach().reshape(-1), torch.int64)
    xl = be.to_device(torch.as_tensor(train_label).detach().reshape(-1), torch.int64)
    hits = be.to_host(be.hit_at_k(idx, ql, xl, [min(int(k), kmax) for k in ks]))
    nq = int(idx.shape[0])
    accs = [float(np.float32(h) / np.float32(nq)) for h in hits]    # .float().mean().item() of a 0/1 vector
    for k, acc in zip(ks, accs):
        print('%dNN acc = %.4f' % (k, acc))
    return accs


def _pdist_forward(be, a, b, dist_metric):
    if dist_metric == 'euclidean':
        return be.distance_matrix(a, b, metric="euclidean")
    ua, _ = be.normalize_rows(a, want_f16=False)
    ub, _ = be.normalize_rows(b, want_f16=False)
    return be.distance_matrix(ua, ub, metric="cosine")


class _PdistFn(torch.autograd.Function):
    """Forward through the C ABI (no [A, B, D] intermediate); backward in closed form with torch ops - the loss
    strategies 'noise_contrastive' / 'all_semi_hard' differentiate through the matrix (triplet_loss.py:100, :122)."""

    @staticmethod
    def forward(ctx, a, b, dist_metric, be):
        out = _pdist_forward(be, a.detach(), b.detach(), dist_metric)
        ctx.metric = dist_metric
        ctx.save_for_backward(a.detach(), b.detach(), out)
        return out

    @staticmethod
    def backward(ctx, g):
        a, b, out = ctx.saved_tensors
        g = g.to(torch.float32)
        if ctx.metric == 'euclidean':
            # d|a_i - b_j| / da_i = (a_i - b_j) / |a_i - b_j|   (0 where the rows coincide)
            w = torch.where(out > 0, g / out.clamp_min(1e-30), torch.zeros_like(g))
            ga = a * w.sum(1, keepdim=True) - w @ b
            gb = b * w.sum(0).unsqueeze(1) - w.t() @ a
        else:
            # D = 1 - ua ub^T with u = x / max(|x|, tiny); the clip to [0, 2] only acts within rounding of its bounds
            na = a.norm(dim=1, keepdim=True).clamp_min(1e-30)
            nb = b.norm(dim=1, keepdim=True).clamp_min(1e-30)
            ua, ub = a / na, b / nb
            gua, gub = -(g @ ub), -(g.t() @ ua)
            ga = (gua - (gua * ua).sum(1, keepdim=True) * ua) / na
            gb = (gub - (gub * ub).sum(1, keepdim=True) * ub) / nb
        return ga, gb, None, None


def pdist_v2(vector1, vector2, eps=1e-6, dist_metric='cosine', backend=None):
    """loss/triplet_loss.py:438-447: [len(vector1), len(vector2)] distance matrix, cosine (1 - cosine similarity) or
    euclidean, in one launch instead of a Python loop over rows.  Differences from the reference, both below 1e-6:
    cosine distances are clipped to [0, 2] (F.cosine_similarity can return 1 + 1e-8); euclidean does not add `eps` to
    every coordinate difference as F.pairwise_distance does.
    Inputs that require grad (the matrix feeds the loss in 'noise_contrastive' / 'all_semi_hard',
    triplet_loss.py:100, :122) get an autograd node: forward through the kernel, analytic backward.  The mining uses
    (triplet_loss.py:54, :279) run under no_grad and take the plain path."""
    be = backend or _backend.default_backend()
    needs_grad = torch.is_grad_enabled() and any(isinstance(v, torch.Tensor) and v.requires_grad
                                                 for v in (vector1, vector2))
    if needs_grad:
        if not all(isinstance(v, torch.Tensor) and v.is_cuda and v.dtype == torch.float32 for v in (vector1, vector2)):
            raise ValueError("pdist: differentiable inputs must be float32 CUDA tensors (a device / dtype copy would cut "
                             "the autograd graph)")
        return _PdistFn.apply(vector1.contiguous(), vector2.contiguous(), dist_metric, be)
    return _pdist_forward(be, _f32_on_device(be, vector1), _f32_on_device(be, vector2), dist_metric)


def pdist(vectors, eps=1e-6, dist_metric='cosine', backend=None):
    """loss/triplet_loss.py:429-436."""
    return pdist_v2(vectors, vectors, eps, dist_metric, backend=backend)
