"""Does sampling the clocks perturb the timed region?  FINCH steps (bench.py's step_resident) timed with no sampler and
with bench.ClockSampler at several periods, inside ONE process group (diagnostic).
torchrun --nproc-per-node G scripts/sampler_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from video_similarity_search_b200 import synth
from video_similarity_search_b200.backend import CudaBackend
from video_similarity_search_b200.clustering.finch import FINCH
from video_similarity_search_b200.sharded import close_peer_groups, sharded_first_neighbors

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
be = CudaBackend()
x = be.to_device(synth.config("C3"))
search = sharded_first_neighbors(be)
step = lambda: FINCH(x, backend=be, verbose=False, first_neighbors=search)
for _ in range(3):
    step()


def timed(steps=5):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = step()
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=be.device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


for label, period, power in (("no sampler", None, None), ("5 ms, clocks+power+reasons", 0.005, True), ("5 ms, no power", 0.005, False),
                             ("20 ms, no power", 0.02, False), ("100 ms, clocks+power+reasons", 0.1, True), ("no sampler again", None, None)):
    sampler = None
    if period is not None and rank == 0:
        sampler = bench.ClockSampler(local, period_s=period, power=power)
        sampler.launch()
    dist.barrier()
    res = []
    for _ in range(3):
        if sampler:
            sampler.t0 = time.time()
        res.append(timed())
    info = sampler.stop() if sampler else None
    if rank == 0:
        print("%-32s ms per step %s  %s" % (label, " ".join("%.2f" % v for v in res), "" if info is None else "samples %s" % info.get("samples")), flush=True)
close_peer_groups()
dist.destroy_process_group()
