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
    if seed == 4:
        lt, lp = np.arange(n), np.arange(n)[::-1].copy()
    # float64 sums reduced in a different (fixed) order than numpy's, lgamma from CUDA's libm: 1e-12 / 1e-10 absolute
    assert metrics.mutual_info_score(lt, lp, backend=be) == pytest.approx(mo.mutual_info_score(lt, lp), abs=1e-12)
    assert metrics.normalized_mutual_info_score(lt, lp, backend=be) == pytest.approx(mo.normalized_mutual_info_score(lt, lp), abs=1e-12)
    assert metrics.adjusted_mutual_info_score(lt, lp, backend=be) == pytest.approx(mo.adjusted_mutual_info_score(lt, lp), abs=1e-10)
    # run-to-run reproducible (fixed reduction order)
    assert metrics.cluster_scores(lt, lp, backend=be) == metrics.cluster_scores(torch.from_numpy(lt).cuda(), torch.from_numpy(lp).cuda(), backend=be)


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
