"""world_size-2 gloo run of the row-sharded level-0 search (sharded.py) with the numpy stand-in backend:
shard ranges, padding, the all-gather of ids + distances, and identical partitions on every rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.golden.make_golden import CASES, make_input


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, golden_dir, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests.fake_backend import FakeBackend
        from video_similarity_search_b200.sharded import FINCH_sharded, shard_range, sharded_first_neighbors
        be = FakeBackend()
        x = make_input(CASES[name])
        n = len(x)
        r0, r1, per = shard_range(n, rank, world)
        assert per * world >= n and 0 <= r0 <= r1 <= n
        nn, d, _ = sharded_first_neighbors(be, triangle=False)(be.to_device(x, torch.float32))
        full_nn, full_d, _ = FakeBackend().first_neighbors(torch.from_numpy(x))
        assert torch.equal(nn, full_nn) and torch.equal(d, full_d)
        # only this rank's shard was searched locally
        assert [c[3] for c in be.calls if c[0] == "first_neighbors"] == [(r0, r1)]
        # triangle parts: every rank screens every world-th block pair, one all-reduce (MIN) of the keys merges them
        be2 = FakeBackend()
        nn2, d2, _ = sharded_first_neighbors(be2)(be2.to_device(x, torch.float32))
        assert torch.equal(nn2, full_nn) and torch.equal(d2, full_d)
        assert [c[2:] for c in be2.calls if c[0] == "first_neighbors_part"] == [(rank, world)]
        assert not [c for c in be2.calls if c[0] == "first_neighbors"]
        # phase 1 of the two-phase search: every rank received every row's best through the all-reduce MAX
        assert be2.exchanged_bests.shape[0] == n
        # an incomplete part (candidate log overflow on one rank) sends every rank to the row-sharded search
        be3 = FakeBackend()
        part = be3.first_neighbors_part

        def overflowing(mat, p, ps, reduce_max=None):
            keys, unit = part(mat, p, ps, reduce_max=reduce_max)
            if p == ps - 1:
                keys[-1] = 0
            return keys, unit
        be3.first_neighbors_part = overflowing
        nn3, d3, _ = sharded_first_neighbors(be3)(be3.to_device(x, torch.float32))
        assert torch.equal(nn3, full_nn) and torch.equal(d3, full_d)
        assert [c[3] for c in be3.calls if c[0] == "first_neighbors"] == [(r0, r1)]
        # query-sharded retrieval top-k: identical to the unsharded search on every rank (ragged last shard included)
        from video_similarity_search_b200.sharded import topk_neighbors_sharded
        xt = be.to_device(x, torch.float32)
        qt = xt[: max(7, n // 3)].clone()
        for same, qq in ((False, qt), (True, xt)):
            gi, gd = topk_neighbors_sharded(qq, xt, 5, same=same, backend=be)
            ei, ed = FakeBackend().topk_neighbors(qq, xt, 5, same=same)
            assert torch.equal(gi, ei) and torch.equal(gd, ed)
        # host matrix in: every rank uploads its 1 / world of the rows, one all-gather assembles the matrix
        from video_similarity_search_b200.sharded import upload_replicated
        up = upload_replicated(x, backend=FakeBackend())
        assert up.shape == x.shape and torch.equal(up, torch.from_numpy(x))
        assert torch.equal(upload_replicated(torch.from_numpy(x.astype(np.float64)), backend=FakeBackend()), torch.from_numpy(x))
        c, num_clust, _ = FINCH_sharded(x, backend=FakeBackend(), verbose=False)
        np.save(os.path.join(out_dir, "c_rank%d.npy" % rank), c)
        g = np.load(os.path.join(golden_dir, name + ".npz"))
        assert num_clust == g["num_clust"].tolist() and np.array_equal(c, g["c"])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,world", [("gmm_1200x64", 2), ("gmm_777x200_odd", 3)])
def test_sharded_first_neighbors_and_finch_under_gloo(tmp_path, golden_dir, name, world):
    mp.spawn(_worker, args=(world, _free_port(), name, golden_dir, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / ("c_rank%d.npy" % r)) for r in range(world)]
    assert all(np.array_equal(parts[0], p) for p in parts[1:])
