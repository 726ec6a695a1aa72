"""-m "not gpu": host logic of the section-8f widenings (embedding hand-off, vid_clusters hand-over, NMI / AMI
wrappers) with the numpy stand-in backend, against the oracle restatements / scikit-learn."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import metrics_oracle as mo
from tests.fake_backend import FakeBackend
from video_similarity_search_b200 import cluster_io, metrics
from video_similarity_search_b200.handoff import EmbeddingCollector


def test_unshuffled_assignments_last_occurrence_wins_and_text_format(tmp_path):
    rng = np.random.default_rng(0)
    n_data = 1000
    idxs = np.concatenate([rng.permutation(n_data)[:900], rng.integers(0, n_data, 124)])   # gaps and repeats
    labels = rng.integers(0, 57, len(idxs)).astype(np.int32)
    got = cluster_io.unshuffled_assignments(labels, idxs.tolist(), n_data, backend=FakeBackend())
    want = mo.unshuffled_assignments(labels.tolist(), idxs.tolist(), n_data)
    assert [None if g == cluster_io.UNASSIGNED else int(g) for g in got] == want
    # the text file is byte-identical to the reference writer's, 'None' lines included
    cluster_io.write_vid_clusters(tmp_path / "a.txt", got)
    mo.write_vid_clusters(tmp_path / "b.txt", want)
    assert (tmp_path / "a.txt").read_bytes() == (tmp_path / "b.txt").read_bytes()
    with pytest.raises(ValueError):
        cluster_io.read_cluster_labels(str(tmp_path / "a.txt"))     # int('None'), as in the reference reader
    with pytest.raises(IndexError):
        cluster_io.unshuffled_assignments(labels[:3], [0, 1, n_data], n_data, backend=FakeBackend())
    # complete assignment: text and .npy round trips, reference reader agrees
    full_idx = rng.permutation(n_data)
    full_lab = rng.integers(0, 57, n_data).astype(np.int32)
    full = cluster_io.unshuffled_assignments(full_lab, full_idx, n_data, backend=FakeBackend())
    cluster_io.write_vid_clusters(tmp_path / "c.txt", full)
    assert cluster_io.read_cluster_labels(str(tmp_path / "c.txt")) == mo.read_cluster_labels(tmp_path / "c.txt") == full.tolist()
    cluster_io.save_cluster_labels_npy(tmp_path / "c.npy", full)
    assert np.array_equal(cluster_io.load_cluster_labels(str(tmp_path / "c.npy")), full)
    assert np.array_equal(cluster_io.load_cluster_labels(str(tmp_path / "c.txt")), full)
    assert cluster_io.read_cluster_labels(None) is None


@pytest.mark.parametrize("n,r,c,seed", [(500, 7, 40, 0), (3000, 30, 300, 1), (64, 1, 1, 2), (200, 1, 9, 3), (200, 200, 200, 4)])
def test_nmi_ami_wrappers_match_sklearn(n, r, c, seed):
    rng = np.random.default_rng(seed)
    lt = rng.integers(0, r, n) * 3 + 5                  # arbitrary (non-dense) label values
    lp = (lt // 3 + rng.integers(0, max(1, c // max(r, 1)), n)) % c
    if seed == 4:       # a perfect match under a relabelling: NMI = AMI = 1
        lt = rng.integers(0, 20, n)
        lp = (lt * 7 + 3) % 20
    be = FakeBackend()
    assert metrics.normalized_mutual_info_score(lt, lp, backend=be) == pytest.approx(mo.normalized_mutual_info_score(lt, lp), abs=1e-12)
    assert metrics.adjusted_mutual_info_score(lt, lp, backend=be) == pytest.approx(mo.adjusted_mutual_info_score(lt, lp), abs=1e-10)
    assert metrics.mutual_info_score(torch.from_numpy(lt), torch.from_numpy(lp), backend=be) == pytest.approx(mo.mutual_info_score(lt, lp), abs=1e-12)


def test_embedding_collector_single_process_keeps_rows_in_arrival_order():
    col = EmbeddingCollector(10, 4, device="cpu")            # capacity too small on purpose: grows
    rng = np.random.default_rng(0)
    batches = [(torch.from_numpy(rng.standard_normal((b, 4)).astype(np.float32)), torch.arange(b) % 3, torch.arange(b) + 100 * k)
               for k, b in enumerate([6, 6, 5])]
    for e, t, i in batches:
        col.append(e, t, i)
    emb, labels, idxs = col.finish()
    assert torch.equal(emb, torch.cat([b[0] for b in batches]))
    assert labels == torch.cat([b[1] for b in batches]).tolist() and idxs == torch.cat([b[2] for b in batches]).tolist()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _collector_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        col = EmbeddingCollector(8, 3, device="cpu")
        for k in range(3):                                   # batch k of this rank: value encodes (batch, rank, row)
            b = 4
            e = torch.tensor([[k, rank, r] for r in range(b)], dtype=torch.float32)
            col.append(e, torch.full((b,), rank), torch.arange(b) + 10 * k + 100 * rank)
        emb, labels, idxs = col.finish()
        # evaluate.py:189-201 order: batch by batch, ranks concatenated inside a batch
        want = torch.tensor([[k, rk, r] for k in range(3) for rk in range(world) for r in range(4)], dtype=torch.float32)
        assert torch.equal(emb, want)
        assert labels == [rk for k in range(3) for rk in range(world) for r in range(4)]
        assert idxs == [r + 10 * k + 100 * rk for k in range(3) for rk in range(world) for r in range(4)]
        torch.save(emb, os.path.join(out_dir, "emb%d.pt" % rank))
    finally:
        dist.destroy_process_group()


def test_embedding_collector_gathers_batches_rank_major_under_gloo(tmp_path):
    mp.spawn(_collector_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert torch.equal(torch.load(tmp_path / "emb0.pt"), torch.load(tmp_path / "emb1.pt"))


def _hard_retrieval_case(n_train=3000, n_test=500, d=128, k=40, seed=0):
    """Overlapping mixture: kNN accuracy well below 1, so the hit counts say something."""
    rng = np.random.default_rng(seed)
    centres = 0.22 * rng.standard_normal((k, d))
    ytr, yte = rng.integers(0, k, n_train), rng.integers(0, k, n_test)
    tr = (centres[ytr] + rng.standard_normal((n_train, d)) + 0.5).astype(np.float32)     # + 0.5: centring matters
    te = (centres[yte] + rng.standard_normal((n_test, d)) + 0.5).astype(np.float32)
    return tr, ytr, te, yte


def test_coclr_retrieval_wrapper_matches_torch_restatement(capsys):
    from oracle import retrieval_oracle as ro
    from video_similarity_search_b200 import coclr_retrieval as cr
    tr, ytr, te, yte = _hard_retrieval_case()
    accs = cr.nn_retrieval_accuracy(torch.from_numpy(te), torch.from_numpy(yte), torch.from_numpy(tr), torch.from_numpy(ytr),
                                    backend=FakeBackend())
    want, _ = ro.coclr_nn_accuracy(te, yte, tr, ytr)
    assert 0.05 < want[0] < 0.9 and accs == want
    assert capsys.readouterr().out.splitlines()[0] == '1NN acc = %.4f' % want[0]
    for metric, tol in (("cosine", 1e-6), ("euclidean", 1e-4)):
        got = cr.pdist_v2(te[:40], tr[:60], 1e-6, metric, backend=FakeBackend()).numpy()
        np.testing.assert_allclose(got, ro.pdist_v2(te[:40], tr[:60], 1e-6, metric), rtol=0, atol=tol)
