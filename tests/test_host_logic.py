"""Host-side logic on CPU with the numpy stand-in backend (tests/fake_backend.py): the FINCH level
loop / exit rules / min_sim mode switch / req_clust refinement of clustering/finch.py, fit_cluster,
the mask wrappers and the retrieval wrappers - checked against the reference's golden fixtures and
the oracle.  The CUDA kernels themselves are checked by the -m gpu tests."""
import io
import json
import os
import types
from contextlib import redirect_stdout

import numpy as np
import pytest
import torch

from oracle import finch_oracle as fo
from oracle import masks_oracle as mo
from oracle import retrieval_oracle as ro
from tests.fake_backend import FakeBackend
from tests.golden.make_golden import CASES, make_input
from video_similarity_search_b200 import evaluate as ev
from video_similarity_search_b200 import iic_retrieve_clips as iic
from video_similarity_search_b200.clustering import cluster_masks as cm
from video_similarity_search_b200.clustering import finch as fm

SMALL = [k for k in CASES if k not in ("c1_9537x512",)]


@pytest.mark.parametrize("name", SMALL)
def test_finch_host_loop_matches_reference_golden(golden_dir, name, monkeypatch):
    case = CASES[name]
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    x = make_input(case)
    kw = dict(ensure_early_exit=case.get("ensure_early_exit", True), req_clust=case.get("req_clust"), verbose=False)
    if case.get("use_initial_rank"):
        kw["initial_rank"] = g["initial_rank"]
    if case.get("flann_threshold") is not None:
        monkeypatch.setattr(fm, "FLANN_THRESHOLD", case["flann_threshold"])
    c, num_clust, req_c = fm.FINCH(x, backend=FakeBackend(), **kw)
    assert num_clust == g["num_clust"].tolist()
    assert c.dtype == np.int32 and c.shape == g["c"].shape and np.array_equal(c, g["c"])
    if g["req_c"].size:
        assert np.array_equal(req_c, g["req_c"])
    else:
        assert req_c is None


def test_finch_prints_partitions_like_reference():
    x = make_input(CASES["gmm_1200x64"])
    buf = io.StringIO()
    with redirect_stdout(buf):
        _, num_clust, _ = fm.FINCH(x, backend=FakeBackend())
    lines = buf.getvalue().strip().splitlines()
    assert lines == ["Partition %d: %d clusters" % (i, n) for i, n in enumerate(num_clust)]


def test_finch_rejects_other_distances_and_bad_rank():
    x = make_input(CASES["iid_600x32"])
    with pytest.raises(NotImplementedError):
        fm.FINCH(x, distance="euclidean", backend=FakeBackend())
    with pytest.raises(ValueError):
        fm.FINCH(x, initial_rank=np.zeros(5, dtype=np.int64), backend=FakeBackend())


def test_finch_accepts_torch_and_float64_input():
    x = make_input(CASES["gmm_777x200_odd"])
    a, na, _ = fm.FINCH(torch.from_numpy(x.astype(np.float64)), backend=FakeBackend(), verbose=False)
    b, nb, _ = fm.FINCH(x, backend=FakeBackend(), verbose=False)
    assert na == nb and np.array_equal(a, b)


def test_level_dtypes_follow_reference():
    """float32 at level 0, float64 centroids at levels >= 1."""
    be = FakeBackend()
    fm.FINCH(make_input(CASES["gmm_1200x64"]), backend=be, verbose=False)
    kinds = [c[2] for c in be.calls if c[0] == "first_neighbors"]
    assert kinds[0] == "torch.float32" and all(k == "torch.float64" for k in kinds[1:])


def test_fit_cluster_finch_partition(monkeypatch):
    from video_similarity_search_b200 import backend as bk
    monkeypatch.setattr(bk, "_default", FakeBackend())
    x = make_input(CASES["gmm_1200x64"])
    with redirect_stdout(io.StringIO()):
        lab0 = cm.fit_cluster(torch.from_numpy(x), method="finch", finch_partition=0)
        lab1 = cm.fit_cluster(torch.from_numpy(x), method="finch", finch_partition=1)
    exp = fo.fit_cluster_finch(x, 0)
    assert lab0.dtype == np.int32 and np.array_equal(lab0, exp)
    assert np.array_equal(lab1, fo.fit_cluster_finch(x, 1))
    with pytest.raises(AssertionError):
        cm.fit_cluster(torch.from_numpy(x), method="nope")
    with pytest.raises(NotImplementedError), redirect_stdout(io.StringIO()):
        cm.fit_cluster(torch.from_numpy(x), method="kmeans")


def test_mask_wrappers_match_call_sites():
    be = FakeBackend()
    rng = np.random.default_rng(0)
    k_label, queue = rng.integers(0, 50, 37), rng.integers(0, 50, 301)
    assert np.array_equal(cm.queue_positive_mask(k_label, queue, backend=be).numpy(),
                          mo.queue_positive_mask(k_label, queue))
    labels = rng.integers(0, 9, 64)
    uniq, pos, neg = mo.in_batch_masks(labels)
    assert np.array_equal(cm.positive_mask(uniq, labels, backend=be).numpy(), pos)
    assert np.array_equal(cm.negative_mask(uniq, labels, backend=be).numpy(), neg)
    bits = cm.positive_mask_bits(labels, backend=be).numpy().view(np.uint32)
    full = labels[:, None] == labels[None, :]
    for j in range(64):
        assert np.array_equal(((bits[:, j // 32] >> (j % 32)) & 1).astype(bool), full[:, j])
    table = cm.label_to_indices(labels * 3 + 1, backend=be)
    exp = mo.label_to_indices(labels * 3 + 1)
    assert table.keys() == exp.keys() and all(np.array_equal(table[k], exp[k]) for k in exp)


def test_evaluate_wrappers_match_oracle():
    be = FakeBackend()
    rng = np.random.default_rng(1)
    x = rng.standard_normal((300, 48)).astype(np.float32)
    q = rng.standard_normal((70, 48)).astype(np.float32)
    xl, ql = rng.integers(0, 6, 300), rng.integers(0, 6, 70)
    dm = ev.get_distance_matrix(q, x, backend=be)
    ref = ro.distance_matrix(q, x)
    assert dm.shape == ref.shape and dm.dtype == ref.dtype
    np.testing.assert_allclose(np.asarray(dm), ref, rtol=0, atol=2e-6)
    assert np.array_equal(ev.get_closest_data_mat(dm, 20), ro.closest_data_mat(ref, 20))
    np.testing.assert_array_equal(ev.get_topk_acc(dm, ql.tolist(), xl.tolist()), ro.topk_acc(ref, ql.tolist(), xl.tolist()))
    # one-set form: diagonal is +inf, neighbours exclude self (evaluate.py:221-222)
    dm1 = ev.get_distance_matrix(x, backend=be)
    ref1 = ro.distance_matrix(x)
    assert np.isinf(np.asarray(dm1)[5, 5])
    assert np.array_equal(ev.get_closest_data_mat(dm1, 5), ro.closest_data_mat(ref1, 5))
    assert np.array_equal(ev.get_closest_data(dm1, 7, 5), ro.closest_data(ref1, 7, 5))
    # dense ndarray input (the reference's own calling convention) and euclidean
    assert np.array_equal(ev.get_closest_data_mat(ref, 10, backend=be), ro.closest_data_mat(ref, 10))
    dme = ev.get_distance_matrix(q.astype(np.float64), x.astype(np.float64), "euclidean", backend=be)
    np.testing.assert_allclose(np.asarray(dme), ro.distance_matrix(q.astype(np.float64), x.astype(np.float64), "euclidean"),
                               rtol=1e-9, atol=1e-9)
    # lazy=False: the ndarray itself, as the reference returns it
    dense = ev.get_distance_matrix(q, x, backend=be, lazy=False)
    assert isinstance(dense, np.ndarray) and dense.dtype == ref.dtype
    np.testing.assert_allclose(dense, ref, rtol=0, atol=2e-6)
    np.testing.assert_array_equal(ev.get_topk_acc(dense, ql.tolist(), xl.tolist(), backend=be), ro.topk_acc(ref, ql.tolist(), xl.tolist()))
    with pytest.raises(AssertionError):
        ev.get_distance_matrix(q, x, "manhattan", backend=be)


def test_iic_topk_retrieval_files(tmp_path):
    be = FakeBackend()
    rng = np.random.default_rng(2)
    centres = rng.standard_normal((7, 32))
    ytr, yte = rng.integers(0, 7, 120), rng.integers(0, 7, 40)
    xtr = centres[ytr][:, None, :] + rng.standard_normal((120, 10, 32))
    xte = centres[yte][:, None, :] + rng.standard_normal((40, 10, 32))
    np.save(tmp_path / "train_feature.npy", xtr)
    np.save(tmp_path / "train_class.npy", np.repeat(ytr[:, None], 10, 1))
    np.save(tmp_path / "test_feature.npy", xte)
    np.save(tmp_path / "test_class.npy", np.repeat(yte[:, None], 10, 1))
    buf = io.StringIO()
    with redirect_stdout(buf):
        got = iic.topk_retrieval(types.SimpleNamespace(feature_dir=str(tmp_path)), backend=be)
    exp, _, _ = ro.topk_retrieval_counts(xtr, np.repeat(ytr[:, None], 10, 1), xte, np.repeat(yte[:, None], 10, 1))
    assert got == exp
    assert json.load(open(tmp_path / "topk_correct.json")) == {str(k): v for k, v in exp.items()}
    assert buf.getvalue().splitlines()[0] == "Load local .npy files."
    assert "Top-50, correct = " in buf.getvalue()


def test_pinned_result_pool_never_reuses_a_buffer_the_caller_still_holds():
    """backend.PinnedResultPool (results are delivered in page-locked buffers without a host-side copy): a buffer goes
    back into circulation only after the array handed out AND every view of it are gone."""
    import gc
    from video_similarity_search_b200.backend import PinnedResultPool
    allocated = []

    def alloc(nbytes):
        allocated.append(nbytes)
        return torch.empty(nbytes, dtype=torch.uint8)          # (pageable stand-in: no CUDA on this box)

    pool = PinnedResultPool(keep=2, alloc=alloc)
    raw, addr_a = pool.take(800)
    c = raw[:800].view(np.int32).reshape(100, 2)               # what finch_host returns: a view of the buffer
    c[:] = 7
    del raw
    other, addr_b = pool.take(800)                             # the first result is still alive: another buffer
    assert addr_b != addr_a and len(allocated) == 2
    other[:] = 0
    assert (c == 7).all()
    column = c[:, 0]                                           # a view of a view keeps the buffer out of circulation
    del c
    gc.collect()
    third, addr_c = pool.take(400)
    assert addr_c not in (addr_a, addr_b) and len(allocated) == 3
    assert (column == 7).all()
    del column, other
    gc.collect()
    again, addr_d = pool.take(400)                             # both early buffers are idle now: one of them is reused
    assert addr_d in (addr_a, addr_b) and len(allocated) == 3
    big, _ = pool.take(1 << 23)                                # larger than anything kept: a new buffer, idle small ones go
    assert len(allocated) == 4 and big.nbytes == 1 << 23
    assert len(pool._entries) <= 4
    del third, again, big
    gc.collect()
    hoard = []                                                 # a caller that keeps every result: page-locked memory stays
    for i in range(5):                                         # bounded - from `keep` + 1 results on, ordinary copies
        raw, _ = pool.take(64)
        raw[:] = i
        hoard.append(pool.deliver(raw[:64].view(np.int32)))
        del raw
    gc.collect()
    assert len(pool._entries) <= 4 and all((h == i * 0x01010101).all() for i, h in enumerate(hoard))
    assert sum(1 for h in hoard if h.flags.owndata) >= 2
    del hoard
    gc.collect()
    for size in (1 << 23) + 1, (1 << 23) + 2, (1 << 23) + 3:  # ever larger requests with everything idle: the pool stays small
        held, _ = pool.take(size)
        del held
        gc.collect()
    assert len(pool._entries) <= 2
