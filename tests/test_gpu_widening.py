"""-m gpu: the section-8f widenings through the C ABI against their oracles - label hand-over (exact), cluster-quality
scores (float64, tolerance stated per assert), device-resident embedding hand-off feeding FINCH."""
import numpy as np
import pytest
import torch

from oracle import finch_oracle as fo
from oracle import metrics_oracle as mo
from video_similarity_search_b200 import cluster_io, metrics, synth
from video_similarity_search_b200.handoff import EmbeddingCollector

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def be():
    from video_similarity_search_b200.backend import CudaBackend
    return CudaBackend()


@pytest.mark.parametrize("n,n_data,seed", [(1, 1, 0), (1000, 900, 1), (240000, 240000, 2), (250000, 240000, 3)])
def test_scatter_last_wins_bit_exact(be, n, n_data, seed):
    rng = np.random.default_rng(seed)
    idxs = rng.integers(0, n_data, n) if seed != 2 else rng.permutation(n_data)
    labels = rng.integers(0, 21436, n).astype(np.int32)
    got = cluster_io.unshuffled_assignments(torch.from_numpy(labels), torch.from_numpy(idxs), n_data, backend=be)
    want = mo.unshuffled_assignments(labels.tolist(), idxs.tolist(), n_data)
    assert got.dtype == np.int32
    assert [None if g == cluster_io.UNASSIGNED else int(g) for g in got.tolist()] == want
    with pytest.raises(IndexError):
        cluster_io.unshuffled_assignments(labels[:1], [n_data], n_data, backend=be)


@pytest.mark.parametrize("n,r,c,seed", [(500, 7, 40, 0), (3000, 30, 300, 1), (64, 1, 1, 2), (200, 1, 9, 3), (200, 200, 200, 4),
                                        (100000, 101, 9000, 5)])
def test_nmi_ami_match_sklearn(be, n, r, c, seed):
    rng = np.random.default_rng(seed)
    lt = rng.integers(0, r, n) * 3 + 5
    lp = (lt // 3 * max(1, c // r) + rng.integers(0, max(1, c // max(r, 1)), n)) % c
    if seed == 4:       # a perfect match under a relabelling: NMI = AMI = 1
        lt = rng.integers(0, 20, n)
        lp = (lt * 7 + 3) % 20
    # float64 sums reduced in a different (fixed) order than numpy's, lgamma from CUDA's libm: 1e-12 / 1e-10 absolute
    assert metrics.mutual_info_score(lt, lp, backend=be) == pytest.approx(mo.mutual_info_score(lt, lp), abs=1e-12)
    assert metrics.normalized_mutual_info_score(lt, lp, backend=be) == pytest.approx(mo.normalized_mutual_info_score(lt, lp), abs=1e-12)
    assert metrics.adjusted_mutual_info_score(lt, lp, backend=be) == pytest.approx(mo.adjusted_mutual_info_score(lt, lp), abs=1e-10)
    # run-to-run reproducible (fixed reduction order)
    assert metrics.cluster_scores(lt, lp, backend=be) == metrics.cluster_scores(torch.from_numpy(lt).cuda(), torch.from_numpy(lp).cuda(), backend=be)


def test_nmi_of_all_singletons_is_one(be):
    """Every row its own class on both sides: NMI = 1.  (AMI is 0 / 0 there - mi, emi and the entropies all equal
    log n - and sklearn's own value is rounding noise over eps; not pinned.)"""
    lt = np.arange(300)
    assert metrics.normalized_mutual_info_score(lt, lt[::-1].copy(), backend=be) == pytest.approx(1.0, abs=1e-12)


def test_nmi_ami_kinetics_size_on_finch_labels(be):
    """online_train.py:633-642 at BASELINE config 3 size: true labels = mixture component (400), predicted = the level-0
    FINCH partition of the 240 000 x 512 embeddings (~21 k clusters)."""
    from video_similarity_search_b200.clustering.finch import FINCH
    x, lab, _ = synth.gaussian_mixture(240000, 512, 400, 0, return_labels=True)
    c, num, _ = FINCH(torch.from_numpy(x).cuda(), backend=be, verbose=False)
    pred = c[:, 0]
    assert metrics.normalized_mutual_info_score(lab, pred, backend=be) == pytest.approx(mo.normalized_mutual_info_score(lab, pred), abs=1e-12)
    assert metrics.adjusted_mutual_info_score(lab, pred, backend=be) == pytest.approx(mo.adjusted_mutual_info_score(lab, pred), abs=1e-10)


def test_cluster_metrics_rejects_oversized_contingency(be):
    from video_similarity_search_b200 import _lib
    n = 70000
    lt = torch.arange(n, dtype=torch.int32, device="cuda")
    with pytest.raises(_lib.SlicError, match="status -3"):
        be.cluster_metrics(lt, lt, n, n)


def test_device_resident_handoff_feeds_finch(be):
    """evaluate.py:170-201 + cluster_masks.py:80 without the host round trip: batches appended on the device, the
    collected matrix goes straight into FINCH; partition equals the oracle's on the same rows."""
    from video_similarity_search_b200.clustering.cluster_masks import fit_cluster
    x = synth.gaussian_mixture(3000, 128, 30, 7)
    col = EmbeddingCollector(3000, 128)
    for s in range(0, 3000, 256):
        e = torch.from_numpy(x[s:s + 256]).cuda()
        col.append(e, torch.zeros(len(e), dtype=torch.int64), torch.arange(s, s + len(e)))
    emb, labels, idxs = col.finish()
    assert emb.is_cuda and emb.shape == (3000, 128) and idxs == list(range(3000))
    got = fit_cluster(emb, method='finch', finch_partition=0)
    co, no, _ = fo.finch(x)
    assert np.array_equal(got, co[:, 0])
    full = cluster_io.unshuffled_assignments(got, idxs, 3000, backend=be)
    assert np.array_equal(full, co[:, 0])


@pytest.mark.parametrize("n_train,n_test,d", [(3000, 500, 128), (9537, 3783, 512)])
def test_coclr_retrieval_matches_torch_restatement(be, n_train, n_test, d):
    """coclr_classify.py:784-810: kNN accuracies at k = 1, 5, 10, 20, 50 equal the torch restatement's; the top-50 SETS
    agree on every row whose 50th / 51st similarity gap exceeds 1e-6 (float32 summation order decides closer ties)."""
    from oracle import retrieval_oracle as ro
    from tests.test_widening_host import _hard_retrieval_case
    from video_similarity_search_b200 import coclr_retrieval as cr
    tr, ytr, te, yte = _hard_retrieval_case(n_train, n_test, d, 40, 1)
    accs = cr.nn_retrieval_accuracy(te, yte, tr, ytr, backend=be)
    want, sim = ro.coclr_nn_accuracy(te, yte, tr, ytr)
    assert 0.05 < want[0] < 0.9
    assert accs == want
    idx, s = cr.retrieval_topk(te, tr, 50, backend=be)
    idx, s = idx.cpu().numpy(), s.cpu().numpy()
    order = np.argsort(-sim, axis=1)[:, :51]
    top = np.take_along_axis(sim, order, 1)
    clear = (top[:, 49] - top[:, 50]) > 1e-6
    assert clear.mean() > 0.95
    assert all(set(idx[i]) == set(order[i, :50]) for i in np.flatnonzero(clear))
    np.testing.assert_allclose(s, top[:, :50], rtol=0, atol=2e-6)          # similarity scores, float32
    # centring alone: column means are zero to float32 rounding
    c = be.center_columns(torch.from_numpy(tr).cuda()).cpu().numpy()
    np.testing.assert_allclose(c, tr - tr.astype(np.float64).mean(0).astype(np.float32), rtol=0, atol=0)


def test_pdist_matches_torch_restatement(be):
    from oracle import retrieval_oracle as ro
    from video_similarity_search_b200 import coclr_retrieval as cr
    rng = np.random.default_rng(3)
    a, b = rng.standard_normal((208, 512)).astype(np.float32), rng.standard_normal((77, 512)).astype(np.float32)
    for metric, tol in (("cosine", 1e-6), ("euclidean", 1e-4)):
        np.testing.assert_allclose(cr.pdist_v2(a, b, 1e-6, metric, backend=be).cpu().numpy(), ro.pdist_v2(a, b, 1e-6, metric),
                                   rtol=0, atol=tol)
        np.testing.assert_allclose(cr.pdist(a, 1e-6, metric, backend=be).cpu().numpy(), ro.pdist_v2(a, a, 1e-6, metric),
                                   rtol=0, atol=tol if metric == "cosine" else 4e-3)   # euclidean self-distances: eps * sqrt(d) in the reference


def test_pdist_backward_matches_torch_autograd(be):
    """pdist feeds the loss in 'noise_contrastive' / 'all_semi_hard' (loss/triplet_loss.py:100, :122): the kernel-backed
    matrix must carry a gradient equal to autograd through the reference's row-by-row formula (:429-447)."""
    import torch.nn.functional as F
    from video_similarity_search_b200 import coclr_retrieval as cr
    g = torch.Generator().manual_seed(5)
    a0, b0 = torch.randn(48, 128, generator=g), torch.randn(33, 128, generator=g)
    w = torch.randn(48, 33, generator=g).cuda()
    for metric in ("cosine", "euclidean"):
        a, b = a0.cuda().requires_grad_(True), b0.cuda().requires_grad_(True)
        out = cr.pdist_v2(a, b, 1e-6, metric, backend=be)
        assert out.requires_grad
        (out * w).sum().backward()
        ar, br = a0.cuda().requires_grad_(True), b0.cuda().requires_grad_(True)
        if metric == "cosine":
            ref = torch.stack([1 - F.cosine_similarity(ar[i].unsqueeze(0), br, dim=1) for i in range(len(ar))])
        else:
            ref = torch.stack([F.pairwise_distance(ar[i].unsqueeze(0), br, eps=0.0) for i in range(len(ar))])
        (ref * w).sum().backward()
        torch.testing.assert_close(a.grad, ar.grad, rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(b.grad, br.grad, rtol=1e-4, atol=1e-5)
        # the same tensor on both sides (pdist): the two gradient paths add up
        v = a0.cuda().requires_grad_(True)
        cr.pdist(v, 1e-6, "cosine", backend=be).pow(2).sum().backward()
        vr = a0.cuda().requires_grad_(True)
        vn = F.normalize(vr, dim=1)
        (1 - vn @ vn.t()).pow(2).sum().backward()
        torch.testing.assert_close(v.grad, vr.grad, rtol=1e-3, atol=1e-5)
    with torch.no_grad():      # the mining call sites (:54, :279): no graph, plain path
        assert not cr.pdist(a0.cuda().requires_grad_(True), backend=be).requires_grad
    with pytest.raises(ValueError):
        cr.pdist_v2(a0.requires_grad_(True), b0, backend=be)     # CPU tensor that needs grad: refuse instead of cutting the graph


@pytest.mark.parametrize("n,lo,hi,seed", [(1, 5, 6, 0), (1000, -50, 50, 1), (240000, 0, 21436, 2), (100000, -2**31, 2**31 - 1, 3),
                                          (5000, 7, 8, 4)])
def test_dense_labels_equal_numpy_unique(be, n, lo, hi, seed):
    """slic_dense_labels = np.unique(labels, return_inverse=True) for int32 labels of either sign (bit-exact)."""
    from video_similarity_search_b200.clustering import cluster_masks as cm
    rng = np.random.default_rng(seed)
    lab = rng.integers(lo, hi, n, dtype=np.int64).astype(np.int32)
    dense, uniq, count = be.dense_labels(torch.from_numpy(lab).cuda())
    eu, einv = np.unique(lab, return_inverse=True)
    assert count == len(eu)
    assert np.array_equal(uniq.cpu().numpy(), eu) and np.array_equal(dense.cpu().numpy(), einv.astype(np.int32))
    if n <= 5000:
        table = cm.label_to_indices(lab, backend=be)
        assert sorted(table) == eu.tolist()
        assert all(np.array_equal(table[v], np.where(lab == v)[0]) for v in eu.tolist())
    with pytest.raises(TypeError):
        metrics.normalized_mutual_info_score(lab.astype(np.float32), lab, backend=be)
