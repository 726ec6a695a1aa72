"""Why does the stand-alone sharded level-0 stage time differently from the same stage inside FINCH?  Per-call CUDA-event
and wall times of search(x) in several orders (diagnostic).  torchrun --nproc-per-node G scripts/nn_only_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
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
for _ in range(3):
    FINCH(x, backend=be, verbose=False, first_neighbors=search)


def per_call(label, fn, reps=6, keep=False):
    dist.barrier(); torch.cuda.synchronize()
    rows = []
    held = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        rows.append((e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3))
        if keep:
            held.append(out)
    print("rank %d %-28s %s" % (rank, label, " ".join("%.2f/%.2f" % r for r in rows)), flush=True)


per_call("search, results dropped", lambda: search(x))
per_call("search, results kept", lambda: search(x), keep=True)
per_call("finch", lambda: FINCH(x, backend=be, verbose=False, first_neighbors=search))
per_call("search again", lambda: search(x))
per_call("normalise only", lambda: be.normalize_rows(x))
per_call("single-GPU search", lambda: be.first_neighbors(x), reps=3)
close_peer_groups()
dist.destroy_process_group()
