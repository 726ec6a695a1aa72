"""-m gpu: every CUDA kernel through the C ABI against the oracle (bit-exact for integer / index work,
stated tolerances for floating point)."""
import os

import numpy as np
import pytest
import torch

from oracle import finch_oracle as fo
from oracle import masks_oracle as mo
from oracle import retrieval_oracle as ro
from tests.fake_backend import FakeBackend
from tests.golden.make_golden import CASES, make_input
from video_similarity_search_b200 import synth

pytestmark = pytest.mark.gpu

TIE_MARGIN_F32 = 2e-6     # rows whose top-1/top-2 cosine-distance gap is below this are
TIE_MARGIN_F64 = 1e-12    # "ties" (float32 GEMM order decides them); they are counted, never ignored silently


@pytest.fixture(scope="module")
def be():
    from video_similarity_search_b200.backend import CudaBackend
    return CudaBackend()


def dev(be, a, dtype=None):
    return be.to_device(a, dtype)


# ---------------------------------------------------------------------------------------------------
# integer primitives through their public users
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,c", [(1, 1), (33, 5), (5000, 700), (4097, 4097), (300000, 21436), (100000, 3),
                                 (70000, 70000),
                                 # one-sweep sort: 17-18 key bits take two 9-bit passes, 19-24 three 8-bit passes,
                                 # 25+ four; tile boundaries (4096 keys) and the scan's (2048 items) at +-1
                                 (300000, 150000), (1000003, 400000), (4096, 2048), (4097, 2049), (8191, 4095),
                                 (2048, 300), (2049, 2049), (600000, 20000000)])
def test_group_by_label_bit_exact(be, n, c):
    rng = np.random.default_rng(n + c)
    lab = rng.integers(0, c, n).astype(np.int32)
    order, off = be.group_by_label(dev(be, lab), c)
    eo, eoff = mo.group_by_label(lab, c)
    assert np.array_equal(order.cpu().numpy(), eo)
    assert np.array_equal(off.cpu().numpy(), eoff)


@pytest.mark.parametrize("n", [1, 2, 1000, 9537, 240000])
def test_components_plain_matches_scipy_numbering(be, n):
    rng = np.random.default_rng(n)
    nn = rng.integers(0, max(n, 1), n).astype(np.int32)
    lab, cnt = be.components(dev(be, nn))
    elab, ecnt = FakeBackend().components(torch.from_numpy(nn))
    assert cnt == ecnt and np.array_equal(lab.cpu().numpy(), elab.numpy())


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_components_filtered_and_min_sim(be, dtype):
    x = synth.gaussian_mixture(1500, 96, 20, 3).astype(dtype)
    fb = FakeBackend()
    nn_f, d_f, unit_f = fb.first_neighbors(torch.from_numpy(x))
    xd = dev(be, x)
    nn, d, unit = be.first_neighbors(xd)
    assert np.array_equal(nn.cpu().numpy(), nn_f.numpy())
    ms = be.min_sim(nn, unit, d)
    ms_f = fb.min_sim(nn_f, unit_f, d_f)
    assert abs(float(ms) - float(ms_f)) <= 1e-6 * float(ms_f)
    for frac in (0.3, 0.6, 1.0):
        thr = float(ms_f) * frac
        lab, cnt = be.components(nn, min_sim=thr, unit=unit, dist=d)
        elab, ecnt = fb.components(nn_f, min_sim=thr, unit=unit_f, dist=d_f)
        assert cnt == ecnt and np.array_equal(lab.cpu().numpy(), elab.numpy()), frac
    i, j = be.closest_link(nn, unit, d)
    assert (i, j) == fb.closest_link(nn_f, unit_f, d_f)


def test_compose_and_segmented_mean(be):
    x = synth.gaussian_mixture(20000, 512, 40, 5)
    rng = np.random.default_rng(0)
    for c in (1, 7, 3000):
        lab = np.unique(rng.integers(0, c, len(x)), return_inverse=True)[1].astype(np.int32)
        k = int(lab.max()) + 1
        m = be.segmented_mean(dev(be, x), dev(be, lab), k).cpu().numpy()
        ref = fo.cluster_means(x, lab)
        assert m.dtype == np.float64 and m.shape == ref.shape
        np.testing.assert_allclose(m, ref, rtol=1e-5, atol=1e-12)          # north_star tolerance: 1e-5 relative
        assert np.max(np.abs(m - ref)) < 1e-11                             # and in fact float64-exact
    prev = rng.integers(0, 50, 1000).astype(np.int32)
    u = rng.integers(0, 9, 50).astype(np.int32)
    assert np.array_equal(be.compose_labels(dev(be, prev), dev(be, u)).cpu().numpy(), u[prev])
    xo = synth.gaussian_mixture(333, 77, 5, 1)                               # odd D: scalar path
    lab = (np.arange(333) % 4).astype(np.int32)
    np.testing.assert_allclose(be.segmented_mean(dev(be, xo), dev(be, lab), 4).cpu().numpy(),
                               fo.cluster_means(xo, lab), rtol=0, atol=1e-11)


def test_copy_to_device_pageable_and_pinned_sources(be):
    """slic_copy_to_device: a pageable source is staged through pinned buffers by several host threads in 32 MB pieces, a
    pinned one goes straight to the DMA engine - the bytes on the device must be the source's either way (sizes around
    the piece and slice boundaries, and a second call that reuses the staging slots)."""
    rng = np.random.default_rng(11)
    for nbytes in (8 << 20, (32 << 20) + 4096, (70 << 20) + 12, 1000):
        src = torch.from_numpy(rng.integers(0, 256, nbytes, dtype=np.uint8))
        for pinned in (False, True):
            host = src.pin_memory() if pinned else src
            dst = torch.empty(nbytes, dtype=torch.uint8, device=be.device)
            be.copy_to_device(dst, host)
            be.copy_to_device(dst, host)
            torch.cuda.synchronize()
            assert torch.equal(dst.cpu(), src), (nbytes, pinned)
    x = rng.standard_normal((40000, 96)).astype(np.float32)          # to_device takes the same route for large host arrays
    assert np.array_equal(be.to_device(x).cpu().numpy(), x)


def test_direct_callers_with_sloppy_labels_get_defined_results(be):
    """the C ABI is public.  Empty clusters of slic_cluster_sums get sum 0 / count 0 / mean NaN instead of
    uninitialised rows; labels outside [0, num) in slic_cluster_metrics are refused (no out-of-bounds write)."""
    x = synth.gaussian_mixture(500, 64, 5, 2)
    lab = (np.arange(500) % 3 * 2).astype(np.int32)                     # clusters 1, 3, 5, 6 are empty
    sums, counts, means = be.cluster_sums(dev(be, x), dev(be, lab), 7)
    assert counts.cpu().numpy().tolist() == [167, 0, 167, 0, 166, 0, 0]
    m = means.cpu().numpy()
    assert np.isnan(m[[1, 3, 5, 6]]).all() and np.all(sums.cpu().numpy()[[1, 3, 5, 6]] == 0)
    np.testing.assert_allclose(m[[0, 2, 4]], fo.cluster_means(x, lab // 2), rtol=0, atol=1e-11)
    t = torch.tensor([0, 1, 2, 3], dtype=torch.int32, device=be.device)
    with pytest.raises(ValueError, match="outside"):
        be.cluster_metrics(t, t, 3, 4)                                    # label 3 >= num_true 3
    with pytest.raises(ValueError, match="2 first-neighbour indices"):
        be.components(torch.tensor([1, 0, 7, -1, 3], dtype=torch.int32, device=be.device))
    lab, cnt = be.components(torch.tensor([1, 0, 3, 2, 3], dtype=torch.int32, device=be.device))
    assert cnt == 2 and lab.cpu().tolist() == [0, 0, 1, 1, 1]
    assert be.cluster_metrics(t, t, 4, 4)[0] > 0


def test_cluster_sums_and_hierarchical_merge(be):
    """slic_cluster_sums / slic_merge_cluster_sums: the level l+1 means formed from the level-l float64 sums equal
    cool_mean over the original rows (finch.py:58-71) to float64 rounding, at every size of cluster."""
    x = synth.gaussian_mixture(30000, 256, 40, 11)
    rng = np.random.default_rng(1)
    lab0 = np.unique(rng.integers(0, 4000, len(x)), return_inverse=True)[1].astype(np.int32)
    c0 = int(lab0.max()) + 1
    sums, counts, means = be.cluster_sums(dev(be, x), dev(be, lab0), c0)
    assert np.array_equal(counts.cpu().numpy(), np.bincount(lab0, minlength=c0))
    np.testing.assert_allclose(means.cpu().numpy(), fo.cluster_means(x, lab0), rtol=0, atol=1e-11)
    cur = lab0
    for c_next in (300, 7, 1):
        u = np.unique(rng.integers(0, c_next, int(cur.max()) + 1), return_inverse=True)[1].astype(np.int32)
        k = int(u.max()) + 1
        sums, counts, means = be.merge_cluster_sums(sums, counts, dev(be, u), k)
        cur = u[cur]
        assert np.array_equal(counts.cpu().numpy(), np.bincount(cur, minlength=k))
        np.testing.assert_allclose(means.cpu().numpy(), fo.cluster_means(x, cur), rtol=0, atol=1e-11)
        np.testing.assert_allclose(sums.cpu().numpy(), fo.cluster_means(x, cur) * np.bincount(cur)[:, None], rtol=1e-12)


def test_label_masks(be):
    rng = np.random.default_rng(4)
    for (na, nb) in [(1, 1), (37, 301), (64, 64), (1024, 65536), (5, 4099)]:
        a, b = rng.integers(0, 97, na), rng.integers(0, 97, nb)
        ad, bd = dev(be, a), dev(be, b)
        full = a[:, None] == b[None, :]
        assert np.array_equal(be.label_mask(ad, bd).cpu().numpy(), full)
        assert np.array_equal(be.label_mask(ad, bd, negate=True).cpu().numpy(), ~full)
        assert np.array_equal(be.label_mask(ad, bd, prepend_ones=True).cpu().numpy(), mo.queue_positive_mask(a, b))
        bits = be.label_mask_bits(ad, bd).cpu().numpy().view(np.uint32)
        unpacked = np.unpackbits(bits.view(np.uint8), axis=1, bitorder="little")[:, :nb].astype(bool)
        assert np.array_equal(unpacked, full)


# ---------------------------------------------------------------------------------------------------
# K1
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_normalize_rows(be, dtype):
    x = synth.gaussian_mixture(1000, 200, 9, 2).astype(dtype)
    x[17] = 0                                                    # zero row: norm replaced by 1
    unit, ub = be.normalize_rows(dev(be, x))
    ref = fo._unit_rows(x)
    tol = 3e-7 if dtype == np.float32 else 1e-15
    np.testing.assert_allclose(unit.cpu().numpy(), ref, rtol=tol, atol=tol)
    assert ub.shape == (1000, 256) and torch.all(ub[:, 200:] == 0)
    # float16 screen operand: relative error 2^-11 in the normal range, absolute 2^-25 in the subnormal range
    np.testing.assert_allclose(ub[:, :200].float().cpu().numpy(), ref, rtol=2 ** -11, atol=2 ** -25)


def _check_nn(nn, d, x, margin):
    """nn/d against the oracle's blocked exact search; rows inside the tie margin are excluded and counted."""
    enn, ed, gap = fo.first_neighbors_blocked(x)
    clear = gap > margin
    assert clear.mean() > 0.99
    assert np.array_equal(nn[clear], enn[clear])
    np.testing.assert_allclose(d[clear], ed[clear], rtol=1e-5, atol=1e-6)
    return int((~clear).sum())


@pytest.mark.parametrize("name", ["gmm_1200x64", "iid_600x32", "gmm_777x200_odd", "gmm_3000x128"])
def test_exact_first_neighbors_vs_reference_golden(be, golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    x = make_input(CASES[name])
    unit, _ = be.normalize_rows(dev(be, x), want_f16=False)
    nn, d = be.nn_exact_top1(unit, unit, self_offset=0)
    _check_nn(nn.cpu().numpy(), d.cpu().numpy(), x, TIE_MARGIN_F32)
    _, _, gap = fo.first_neighbors_blocked(x)
    clear = gap > TIE_MARGIN_F32
    assert np.array_equal(nn.cpu().numpy()[clear], g["nn_level0"][clear])      # the reference's own argmin


@pytest.mark.parametrize("nq,n,d", [(128, 256, 64), (100, 300, 64), (257, 1000, 128), (384, 2100, 512), (130, 513, 200)])
def test_tensor_core_screen_scores(be, nq, n, d):
    """Raw tcgen05 scores element by element against a float32 product of the same f16 inputs."""
    rng = np.random.default_rng(nq * n)
    dp = (d + 63) // 64 * 64
    q = torch.zeros((nq, dp), dtype=torch.float16, device=be.device)
    x = torch.zeros((n, dp), dtype=torch.float16, device=be.device)
    q[:, :d] = torch.from_numpy(rng.standard_normal((nq, d)).astype(np.float32) / np.sqrt(d)).to(be.device)
    x[:, :d] = torch.from_numpy(rng.standard_normal((n, d)).astype(np.float32) / np.sqrt(d)).to(be.device)
    got = be.screen_scores_debug(q, x).cpu().numpy()
    ref = q.float().cpu().numpy().astype(np.float64) @ x.float().cpu().numpy().astype(np.float64).T
    np.testing.assert_allclose(got, ref, rtol=0, atol=2e-5)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n,d,k,seed", [(3000, 128, 30, 7), (9537, 512, 101, 0), (5000, 200, 10, 9)])
def test_screened_first_neighbors_equal_exact(be, dtype, n, d, k, seed):
    x = synth.gaussian_mixture(n, d, k, seed).astype(dtype)
    unit, ub = be.normalize_rows(dev(be, x))
    nn_e, d_e = be.nn_exact_top1(unit, unit, self_offset=0)
    nn_s, d_s = be.nn_top1(unit, ub, unit, ub, self_offset=0)
    assert torch.equal(nn_e, nn_s)
    # the two kernels add the float64 products in different orders: equal to rounding of the reference dtype
    np.testing.assert_allclose(d_s.cpu().numpy(), d_e.cpu().numpy(), rtol=0, atol=1.2e-7 if dtype == np.float32 else 1e-14)
    stats = be.last_stats.cpu().numpy()
    assert stats[1] == 0                                          # no row needed the exact finisher
    # row-sharded form (what each rank of the multi-GPU path runs)
    r0, r1 = n // 3, n // 3 + 1000
    nn_p, _ = be.nn_top1(unit[r0:r1], ub[r0:r1], unit, ub, self_offset=r0)
    assert torch.equal(nn_p, nn_e[r0:r1])


def test_screen_overflow_rows_are_finished_exactly(be):
    """80 identical rows: more than 32 columns tie at the top, the candidate list overflows and the exact
    kernel must finish those rows (lowest index wins, as np.argmin does)."""
    x = synth.gaussian_mixture(4000, 64, 8, 21)
    x[100:180] = x[100]
    unit, ub = be.normalize_rows(dev(be, x))
    nn_e, _ = be.nn_exact_top1(unit, unit, self_offset=0)
    nn_s, _ = be.nn_top1(unit, ub, unit, ub, self_offset=0)
    assert torch.equal(nn_e, nn_s)
    assert int(be.last_stats[1]) >= 80
    got = nn_s.cpu().numpy()
    assert got[100] == 101 and np.all(got[101:180] == 100)


def test_c1_first_neighbors_vs_reference_golden(be, golden_dir):
    """BASELINE config 1 through the tensor-core path against the reference's own argmin."""
    g = np.load(os.path.join(golden_dir, "c1_9537x512.npz"))
    x = synth.config("C1")
    nn, d, _ = be.first_neighbors(dev(be, x))
    ties = _check_nn(nn.cpu().numpy(), d.cpu().numpy(), x, TIE_MARGIN_F32)
    _, _, gap = fo.first_neighbors_blocked(x)
    clear = gap > TIE_MARGIN_F32
    assert np.array_equal(nn.cpu().numpy()[clear], g["nn_level0"][clear])
    assert ties <= 10


# ---------------------------------------------------------------------------------------------------
# retrieval
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_distance_matrix_and_topk(be, dtype):
    rng = np.random.default_rng(5)
    x = rng.standard_normal((1500, 96)).astype(dtype)
    q = rng.standard_normal((333, 96)).astype(dtype)
    ux, _ = be.normalize_rows(dev(be, x), want_f16=False)
    uq, _ = be.normalize_rows(dev(be, q), want_f16=False)
    ref = ro.distance_matrix(q, x)
    got = be.distance_matrix(uq, ux).cpu().numpy()
    assert got.dtype == ref.dtype
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=2e-6 if dtype == np.float32 else 1e-13)
    refe = ro.distance_matrix(q, x, "euclidean")
    gote = be.distance_matrix(dev(be, q), dev(be, x), metric="euclidean").cpu().numpy()
    np.testing.assert_allclose(gote, refe, rtol=1e-5, atol=1e-5 if dtype == np.float32 else 1e-12)
    # rows_topk is an exact integer/selection kernel: check it on the SAME matrix the oracle sees
    for k in (1, 5, 50, 333):
        idx, val = be.rows_topk(dev(be, ref), k)
        order = np.lexsort((np.broadcast_to(np.arange(ref.shape[1]), ref.shape), ref), axis=1)[:, :k]
        assert np.array_equal(idx.cpu().numpy(), order)
        assert np.array_equal(val.cpu().numpy(), np.take_along_axis(ref, order, 1))
    # duplicate distances at the boundary: lowest columns win
    tie = np.ones((3, 700), dtype=dtype)
    tie[:, 650:] = 0.5
    idx, _ = be.rows_topk(dev(be, tie), 60)
    assert np.array_equal(idx.cpu().numpy()[0], np.concatenate([np.arange(650, 700), np.arange(0, 10)]))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("nq,n,d,k,same", [(3783, 9537, 512, 50, False), (1000, 5000, 200, 64, False),
                                            (4000, 4000, 128, 20, True), (300, 20000, 96, 1, False),
                                            (2500, 2500, 64, 5, True)])
def test_tensor_core_topk_equals_exact_topk(be, dtype, nq, n, d, k, same):
    """slic_topk_cosine_tc (tcgen05 screen, k-th-best candidate rule, exact re-rank) must return exactly what the
    exact kernels return: same columns in the same order, same distances - bit for bit (both evaluate the
    surviving pairs with float64 accumulation and round to the reference dtype)."""
    x = synth.gaussian_mixture(n, d, 40, 21).astype(dtype)
    q = x if same else synth.gaussian_mixture(nq, d, 40, 22).astype(dtype)
    xd = dev(be, x)
    ux, xb = be.normalize_rows(xd)
    uq, qb = (ux, xb) if same else be.normalize_rows(dev(be, q))
    off = 0 if same else -1
    ei, ev = be.topk_cosine(uq, ux, k, self_offset=off)
    ti, tv = be.topk_cosine(uq, ux, k, self_offset=off, q_f16=qb, x_f16=xb)
    stats = be.last_stats.cpu().numpy()
    ei, ev, ti, tv = ei.cpu().numpy(), ev.cpu().numpy(), ti.cpu().numpy(), tv.cpu().numpy()
    # the two paths accumulate the same float64 products in different orders: a pair of columns whose distances
    # round to the same value in one and to neighbours in the other may swap - compare outside such near-ties
    margin = TIE_MARGIN_F32 if dtype == np.float32 else TIE_MARGIN_F64
    np.testing.assert_allclose(tv, ev, rtol=0, atol=margin)
    bad = np.flatnonzero((ti != ei).any(1))
    assert len(bad) <= max(1, nq // 500)
    for r in bad:                          # a disagreement needs two distances closer than the margin in that row
        assert np.diff(ev[r]).min() <= margin and set(ti[r]) == set(ei[r])
    if same:
        assert not (ti == np.arange(nq)[:, None]).any()
    assert stats[1] <= nq // 50            # rows handed to the exact finisher
    assert stats[0] >= nq * k              # every row re-ranked at least k candidates


def test_tensor_core_topk_overflow_rows_are_finished_exactly(be):
    """Many near-duplicates: far more than RK_MAX columns within eps of the k-th best -> the screen cannot
    settle those rows and the exact kernels finish them; the result is still the exact top-k."""
    rng = np.random.default_rng(3)
    base = rng.standard_normal((4, 64)).astype(np.float32)
    x = (np.repeat(base, 1500, axis=0) + 1e-4 * rng.standard_normal((6000, 64))).astype(np.float32)
    q = (base[rng.integers(0, 4, 200)] + 1e-4 * rng.standard_normal((200, 64))).astype(np.float32)
    ux, xb = be.normalize_rows(dev(be, x))
    uq, qb = be.normalize_rows(dev(be, q))
    ei, ev = be.topk_cosine(uq, ux, 10)
    ti, tv = be.topk_cosine(uq, ux, 10, q_f16=qb, x_f16=xb)
    assert int(be.last_stats.cpu().numpy()[1]) == 200
    assert np.array_equal(ti.cpu().numpy(), ei.cpu().numpy())
    assert np.array_equal(tv.cpu().numpy(), ev.cpu().numpy())


def test_hit_at_k(be):
    rng = np.random.default_rng(6)
    idx = rng.integers(0, 500, (200, 50)).astype(np.int32)
    ql, xl = rng.integers(0, 30, 200), rng.integers(0, 30, 500)
    ks = [1, 5, 10, 20, 50]
    got = be.hit_at_k(dev(be, idx), dev(be, ql), dev(be, xl), ks).cpu().numpy()
    exp = [int((xl[idx[:, :k]] == ql[:, None]).any(1).sum()) for k in ks]
    assert got.tolist() == exp


def test_host_buffer_entry_point(be):
    """slic_first_neighbors_host: the call a reference-side binding makes with numpy arrays."""
    import ctypes
    from video_similarity_search_b200 import _lib
    x = synth.gaussian_mixture(3000, 128, 30, 7)
    nn = np.empty(3000, dtype=np.int32)
    d = np.empty(3000, dtype=np.float32)
    _lib.call("slic_first_neighbors_host", x.ctypes.data_as(ctypes.c_void_p), 3000, 128, 0,
              nn.ctypes.data_as(ctypes.c_void_p), d.ctypes.data_as(ctypes.c_void_p))
    _check_nn(nn, d, x, TIE_MARGIN_F32)


# ---- symmetric self-search (upper-triangular tiles, row + column filters) -----------------------------------
@pytest.mark.parametrize("dtype,n,d,k,seed", [
    (np.float32, 16384, 64, 40, 1),       # exactly at the switch-over, T = 64 tiles = one column chunk
    (np.float32, 20011, 96, 0, 2),        # iid rows (no cluster structure: weak pre-pass thresholds), ragged tail
    (np.float32, 50000, 512, 300, 3),     # several column chunks, the A-resident kernel
    (np.float64, 33000, 128, 100, 4),     # float64 re-rank (FINCH levels >= 1)
    (np.float32, 24000, 640, 50, 5),      # d_pad > 512: the streaming (non-resident) pair kernel
])
def test_symmetric_screen_equals_exact(be, dtype, n, d, k, seed):
    """Self-searches of >= 16 384 rows run the symmetric screen: only tiles on or right of the diagonal are computed
    and each is filtered along rows and along columns.  The result must be that of the exact kernel on every row."""
    if k:
        x = synth.gaussian_mixture(n, d, k, seed).astype(dtype)
    else:
        x = np.random.default_rng(seed).standard_normal((n, d)).astype(dtype)
    xd = be.to_device(x)
    unit, ub = be.normalize_rows(xd)
    idx_tc, dist_tc = be.nn_top1(unit, ub, unit, ub, self_offset=0)
    stats = be.last_stats.cpu().numpy()
    idx_ex, dist_ex = be.nn_exact_top1(unit, unit, self_offset=0)
    assert torch.equal(idx_tc, idx_ex)
    # the two kernels add the float64 products in different orders: equal to rounding of the reference dtype
    np.testing.assert_allclose(dist_tc.cpu().numpy(), dist_ex.cpu().numpy(), rtol=0, atol=1.2e-7 if dtype == np.float32 else 1e-14)
    assert stats[1] < n // 100          # rows handed to the exact finisher stay rare
    lib = be.lib
    import ctypes
    lib.slic_profile_screen(1)
    be.nn_top1(unit, ub, unit, ub, self_offset=0)
    ms, fl, ex = ctypes.c_float(0), ctypes.c_double(0), ctypes.c_double(0)
    assert lib.slic_last_screen_time(ctypes.byref(ms), ctypes.byref(fl)) == 0
    assert lib.slic_last_screen_exec_flop(ctypes.byref(ex)) == 0
    lib.slic_profile_screen(0)
    assert ex.value < (0.6 if n >= 50000 else 0.8) * fl.value   # about half of the square was computed (plus the pre-pass)


def test_symmetric_screen_dense_cluster_rows_overflow_to_exact(be):
    """Many rows within eps of each other: lists overflow, the exact kernel must take over - same result."""
    rng = np.random.default_rng(9)
    base = rng.standard_normal((1, 64)).astype(np.float32)
    x = np.concatenate([base + 1e-3 * rng.standard_normal((600, 64)).astype(np.float32),
                        rng.standard_normal((17000, 64)).astype(np.float32)])
    xd = be.to_device(x)
    unit, ub = be.normalize_rows(xd)
    idx_tc, dist_tc = be.nn_top1(unit, ub, unit, ub, self_offset=0)
    idx_ex, dist_ex = be.nn_exact_top1(unit, unit, self_offset=0)
    assert torch.equal(idx_tc, idx_ex)
    np.testing.assert_allclose(dist_tc.cpu().numpy(), dist_ex.cpu().numpy(), rtol=0, atol=1.2e-7)
    assert int(be.last_stats[1]) >= 600


@pytest.mark.parametrize("n,d,k,seed,parts", [(16384, 64, 40, 1, 2), (50000, 512, 300, 3, 3), (41003, 200, 0, 6, 8)])
def test_symmetric_screen_parts_merge_to_the_full_search(be, n, d, k, seed, parts):
    """Multi-GPU share of the self-search (slic_nn_top1_sym_part): each part screens every parts-th unit of the
    triangle; the element-wise MIN of the parts' (distance, neighbour) keys - what the all-reduce over the ranks
    computes - must be the exact first neighbour of every row, and every part must be complete."""
    if k:
        x = synth.gaussian_mixture(n, d, k, seed)
    else:
        x = np.random.default_rng(seed).standard_normal((n, d)).astype(np.float32)
    xd = be.to_device(x)
    assert be.supports_triangle_parts(xd)
    merged = None
    for part in range(parts):
        keys, unit = be.first_neighbors_part(xd, part, parts)
        assert int(keys[n]) == 1
        merged = keys if merged is None else torch.minimum(merged, keys)
    nn, dist, complete = be.unpack_neighbor_keys(merged)
    assert complete
    idx_ex, dist_ex = be.nn_exact_top1(unit, unit, self_offset=0)
    assert torch.equal(nn, idx_ex)
    np.testing.assert_allclose(dist.cpu().numpy(), dist_ex.cpu().numpy(), rtol=0, atol=1.2e-7)
    # a single part alone is NOT the answer (it saw a fraction of the pairs) - the merge is doing real work
    if parts > 1:
        nn0, _, _ = be.unpack_neighbor_keys(keys)
        assert not torch.equal(nn0, idx_ex)
    # two-phase form (what sharded.py runs): every part first screens ITS rows against a column sample, the row bests
    # are merged by MAX (the all-reduce), and every part starts its share of the triangle with thresholds for all rows
    from video_similarity_search_b200 import _lib
    unit2, ub = be.normalize_rows(xd)
    bests = None
    for part in range(parts):
        b = torch.empty(n, dtype=torch.int32, device=xd.device)
        _lib.call("slic_sym_row_bests", unit2.data_ptr(), ub.data_ptr(), n, d, ub.shape[1], part, parts, b.data_ptr(), None)
        bests = b if bests is None else torch.maximum(bests, b)
    assert int((bests == torch.iinfo(torch.int32).min).sum()) == 0 and int(bests.min()) > -2139095041   # every row has a finite best
    merged2, logged = None, 0
    for part in range(parts):
        keys, _ = be.first_neighbors_part(xd, part, parts, reduce_max=lambda t: t.copy_(bests))
        assert int(keys[n]) == 1
        logged += int(be.last_stats[0])
        merged2 = keys if merged2 is None else torch.minimum(merged2, keys)
    nn2, dist2, complete2 = be.unpack_neighbor_keys(merged2)
    assert complete2 and torch.equal(nn2, idx_ex) and torch.equal(dist2, dist)
    # parts = 1 is the whole triangle
    keys1, _ = be.first_neighbors_part(xd, 0, 1)
    nn1, _, complete1 = be.unpack_neighbor_keys(keys1)
    assert complete1 and torch.equal(nn1, idx_ex)


def test_symmetric_screen_part_rejects_small_inputs(be):
    from video_similarity_search_b200 import _lib
    xd = be.to_device(synth.gaussian_mixture(4000, 64, 10, 1))
    assert not be.supports_triangle_parts(xd)
    with pytest.raises(_lib.SlicError, match="status -3"):
        be.first_neighbors_part(xd, 0, 2)
