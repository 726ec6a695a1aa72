"""Multi-GPU paths on REAL devices (NVLink peer windows, NCCL): skipped on a box with fewer than two GPUs.

  * one process, all GPUs (slic_comm_create / slic_finch_multi): the unmodified rank-0 call site FINCH(host matrix);
  * one process per GPU (torch.distributed.run, NCCL): the fused peer-window search, the round-1 all-reduce scheme and
    the row-sharded top-k retrieval, each against the single-GPU result on every rank, bit for bit.
Run with: gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from video_similarity_search_b200 import synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


needs_two = pytest.mark.skipif(_gpus() < 2, reason="needs at least two GPUs on the box")


@needs_two
def test_rank0_driven_finch_over_all_gpus_equals_one_gpu():
    """slic_finch_multi behind FINCH(host matrix): same partition as slic_finch_host on one GPU (labels exact), for a
    shape the group shares (N >= 16384), a ragged row count, repeated calls on one group, and the inputs it hands to
    device 0 alone (small N, initial_rank)."""
    from video_similarity_search_b200.backend import CudaBackend
    from video_similarity_search_b200.clustering.finch import FINCH
    be1 = CudaBackend()
    bem = CudaBackend()
    bem.enable_multi_gpu(max_rows=70000)
    try:
        for n, d, k, seed in ((40000, 128, 60, 5), (50003, 64, 40, 6), (3000, 128, 30, 7)):
            x = synth.gaussian_mixture(n, d, k, seed)
            c1, num1, _ = FINCH(x, backend=be1, verbose=False)
            for _ in range(2):
                cm, numm, _ = FINCH(x, backend=bem, verbose=False)
                assert numm == num1 and np.array_equal(cm, c1), (n, numm, num1)
        up, search, total = bem.multi_gpu_timeline()
        assert total >= search > 0
        x = synth.gaussian_mixture(20000, 64, 20, 8)
        nn, _, _ = be1.first_neighbors(be1.to_device(x))
        rank = nn.cpu().numpy().astype(np.int64)
        c1, num1, _ = FINCH(x, initial_rank=rank, backend=be1, verbose=False)
        cm, numm, _ = FINCH(x, initial_rank=rank, backend=bem, verbose=False)
        assert numm == num1 and np.array_equal(cm, c1)
        x = synth.gaussian_mixture(80000, 32, 10, 9)     # more rows than the windows hold: device 0 alone
        c1, num1, _ = FINCH(x, backend=be1, verbose=False)
        cm, numm, _ = FINCH(x, backend=bem, verbose=False)
        assert numm == num1 and np.array_equal(cm, c1)
    finally:
        bem.disable_multi_gpu()


_RANK_SCRIPT = r"""
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from video_similarity_search_b200 import synth
from video_similarity_search_b200.backend import CudaBackend
from video_similarity_search_b200.clustering.finch import FINCH
from video_similarity_search_b200.sharded import FINCH_sharded, close_peer_groups, sharded_first_neighbors, topk_neighbors_sharded
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
be = CudaBackend()
ok = True
for n, d, k, seed in ((40000, 128, 60, 5), (50003, 64, 40, 6)):
    x = be.to_device(synth.gaussian_mixture(n, d, k, seed))
    nn1, d1, _ = be.first_neighbors(x)
    for peer in (True, False):
        search = sharded_first_neighbors(be, peer=peer)
        for _ in range(2):
            nn, dd, _ = search(x)
            ok &= bool(torch.equal(nn, nn1) and torch.equal(dd, d1))
    c1, num1, _ = FINCH(x, backend=be, verbose=False)
    cs, nums, _ = FINCH_sharded(x.cpu().numpy(), backend=be, verbose=False)
    ok &= nums == num1 and bool(np.array_equal(cs, c1))
# retrieval top-k with the query rows sharded
tr, _, te, _ = synth.c2_retrieval()
q, xx = be.to_device(te), be.to_device(tr)
i1, v1 = be.topk_neighbors(q, xx, 50)
i2, v2 = topk_neighbors_sharded(q, xx, 50, backend=be)
ok &= bool(torch.equal(i1, i2) and torch.equal(v1, v2))
i1, v1 = be.topk_neighbors(xx, xx, 5, same=True)
i2, v2 = topk_neighbors_sharded(xx, xx, 5, same=True, backend=be)
ok &= bool(torch.equal(i1, i2) and torch.equal(v1, v2))
# a rank that cannot set up its peer window (no IPC in the container, no peer access): EVERY rank must notice, nobody may
# hang, and the search must fall back to the NCCL scheme with the same result
close_peer_groups()
if rank == 1:
    def broken(max_rows):
        raise RuntimeError("simulated: CUDA IPC not permitted")
    be.comm_window_create = broken
import warnings
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    search = sharded_first_neighbors(be)
    nn, dd, _ = search(x)
    ok &= bool(torch.equal(nn, nn1) and torch.equal(dd, d1))
    cs, nums, _ = FINCH_sharded(x, backend=be, verbose=False)
    ok &= nums == num1 and bool(np.array_equal(cs, c1))
t = torch.tensor([int(ok)], device=be.device)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("RESULT", int(t.item()))
close_peer_groups()
dist.destroy_process_group()
"""


@needs_two
def test_one_process_per_gpu_real_nccl_and_peer_windows(tmp_path):
    """torch.distributed.run, 2 ranks, backend nccl: merged first neighbours (peer windows and the all-reduce scheme) and
    sharded top-k equal the single-GPU results on every rank."""
    script = tmp_path / "ranks.py"
    script.write_text(_RANK_SCRIPT % ROOT)
    env = dict(os.environ)
    env.pop("CUDA_LAUNCH_BLOCKING", None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", str(script)], capture_output=True, text=True, env=env,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    assert "RESULT 1" in r.stdout, (r.stdout[-500:], r.stderr[-2000:])
