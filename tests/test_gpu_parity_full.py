"""Full-population parity at the sizes BASELINE.json quotes the metric on ("do this" #2).

  C3  N = 240 000 x D = 512 (configs[2], Kinetics-400 train size): the tensor-core path against the exact
      float64-accumulating kernel on EVERY row; against the oracle's first neighbours (the reference's arithmetic,
      finch.py:27-29, blocked) on 16 384 rows; and the whole partition against the oracle's hierarchy run from those
      level-0 neighbours.  Rows whose top-1 / top-2 gap lies inside the float32 tie margin are counted and printed.
  C5  N = 1 000 000 x D = 1 024 (configs[4], S3D feature width; generated on the device): the streaming (non-resident)
      symmetric screen against the exact kernel on 65 536 rows spread over the matrix, the oracle on 512 of them, and
      the size-independent properties of the result (no self links, symmetric distances on mutual pairs, partition
      sizes strictly decreasing, labels dense).
"""
import time

import numpy as np
import pytest
import torch

from oracle import finch_oracle as fo
from video_similarity_search_b200 import synth
from video_similarity_search_b200.clustering.finch import FINCH

pytestmark = pytest.mark.gpu

TIE_MARGIN_F32 = 2e-6


@pytest.fixture(scope="module")
def be():
    from video_similarity_search_b200.backend import CudaBackend
    return CudaBackend()


def test_c3_every_first_neighbour_and_the_partition(be):
    n, d, k, seed = synth.CONFIGS["C3"]
    x = synth.gaussian_mixture(n, d, k, seed)
    xd = be.to_device(x)
    nn, dist, unit = be.first_neighbors(xd)                       # tcgen05 screen + exact re-rank
    t0 = time.time()
    nn_ex, dist_ex = be.nn_exact_top1(unit, unit, self_offset=0)  # no screen: float64 accumulation over every pair
    torch.cuda.synchronize()
    mism = int((nn != nn_ex).sum())
    print("C3: tensor-core path vs exact kernel on all %d rows: %d mismatches (%.1f s)" % (n, mism, time.time() - t0))
    assert mism == 0 and torch.equal(nn, nn_ex)
    np.testing.assert_allclose(dist.cpu().numpy(), dist_ex.cpu().numpy(), rtol=0, atol=1.2e-7)
    nn_host = nn.cpu().numpy().astype(np.int64)
    assert not np.any(nn_host == np.arange(n))                    # no row is its own neighbour
    rows = np.linspace(0, n - 1, 16384).astype(np.int64)
    enn, _, gap = fo.first_neighbors_blocked(x, rows=rows)
    clear = gap > TIE_MARGIN_F32
    print("C3: oracle first neighbours on %d rows: %d inside the %.0e tie margin, %d differ in all" %
          (len(rows), int((~clear).sum()), TIE_MARGIN_F32, int((nn_host[rows] != enn).sum())))
    assert (~clear).sum() < 64
    assert np.array_equal(nn_host[rows][clear], enn[clear])
    # the whole hierarchy: one native call on the device matrix; the oracle builds every level from the GPU's level-0
    # neighbours (above 70 000 rows the reference has no dense distances: initial_rank semantics, finch.py:30-38)
    c, num_clust, _ = FINCH(xd, backend=be, verbose=False)
    co, no, _ = fo.finch(x, initial_rank=nn_host)
    print("C3: partitions", num_clust)
    assert num_clust == no
    assert np.array_equal(c, co)
    # and from a host matrix through the pipelined upload (the reference-facing call)
    ch, numh, _ = FINCH(x, backend=be, verbose=False)
    assert numh == num_clust and np.array_equal(ch, c)


def test_c5_first_neighbours_on_65536_rows_and_result_properties(be):
    n, d, k, seed = synth.CONFIGS["C5"]
    g = torch.Generator(device=be.device).manual_seed(seed)
    centres = torch.randn(k, d, device=be.device, generator=g)
    xd = torch.empty(n, d, device=be.device)
    for s in range(0, n, 65536):
        e = min(n, s + 65536)
        lab = torch.randint(0, k, (e - s,), device=be.device, generator=g)
        xd[s:e] = centres[lab] + torch.randn(e - s, d, device=be.device, generator=g)
    nn, dist, unit = be.first_neighbors(xd)
    rows = torch.linspace(0, n - 1, 65536, device=be.device).long().to(torch.int32)
    t0 = time.time()
    nn_ex, dist_ex = be.nn_exact_top1(unit, unit, self_offset=0, q_rows=rows)
    torch.cuda.synchronize()
    mism = int((nn_ex != nn[rows.long()]).sum())
    print("C5: tensor-core path vs exact kernel on %d of %d rows: %d mismatches (%.1f s)" % (len(rows), n, mism, time.time() - t0))
    assert mism == 0
    np.testing.assert_allclose(dist[rows.long()].cpu().numpy(), dist_ex.cpu().numpy(), rtol=0, atol=1.2e-7)
    # oracle (the reference's float32 arithmetic on the host) on 512 of those rows
    x_host = xd.cpu().numpy()
    orows = rows[:: len(rows) // 512].cpu().numpy().astype(np.int64)
    enn, _, gap = fo.first_neighbors_blocked(x_host, rows=orows)
    clear = gap > TIE_MARGIN_F32
    nn_host = nn.cpu().numpy().astype(np.int64)
    print("C5: oracle on %d rows: %d inside the tie margin, %d differ in all" %
          (len(orows), int((~clear).sum()), int((nn_host[orows] != enn).sum())))
    assert np.array_equal(nn_host[orows][clear], enn[clear])
    del x_host
    # size-independent properties of the neighbour array
    idx = torch.arange(n, device=be.device, dtype=torch.int32)
    assert not bool((nn == idx).any())
    mutual = nn[nn.long()] == idx
    assert int(mutual.sum()) > n // 10
    dm = dist[mutual]
    assert torch.equal(dm, dist[nn.long()][mutual])               # d(i, j) == d(j, i) bit for bit on mutual pairs
    # the hierarchy on top of it
    c, num_clust, _ = FINCH(xd, backend=be, verbose=False)
    print("C5: partitions", num_clust)
    assert all(a > b for a, b in zip(num_clust, num_clust[1:])) and num_clust[-1] >= 2
    for lvl, cnt in enumerate(num_clust):
        col = c[:, lvl]
        assert col.min() == 0 and col.max() == cnt - 1
    # every level's clusters are unions of the previous level's (get_merge, finch.py:74-79)
    for lvl in range(1, len(num_clust)):
        pairs = np.unique(c[:, lvl - 1].astype(np.int64) * (num_clust[lvl] + 1) + c[:, lvl])
        assert len(pairs) == num_clust[lvl - 1]
