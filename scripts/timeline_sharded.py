"""CUPTI timeline (torch.profiler) of rank 0's share of one multi-GPU level-0 search and one full sharded FINCH step
(diagnostic; never a bench number).  Launch: python -m torch.distributed.run --nproc-per-node G scripts/timeline_sharded.py
Output: gpurun_out/timeline_sharded_{nn,finch}_G<G>.txt (rank 0) + per-stage CUDA-event times of every rank."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
from video_similarity_search_b200 import synth
from video_similarity_search_b200.backend import CudaBackend
from video_similarity_search_b200.clustering.finch import FINCH
from video_similarity_search_b200.sharded import sharded_first_neighbors

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
be = CudaBackend()
x = be.to_device(synth.config(sys.argv[1] if len(sys.argv) > 1 else "C3"))
search = sharded_first_neighbors(be)
for _ in range(3):
    FINCH(x, backend=be, verbose=False, first_neighbors=search)
torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)


def run(name, fn):
    dist.barrier()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    if rank != 0:
        return
    path = "gpurun_out/timeline_sharded_%s_G%d.json" % (name, world)
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("ph") == "X" and e.get("cat") in
          ("kernel", "gpu_memcpy", "gpu_memset", "cuda_runtime", "cuda_driver")]
    ev.sort(key=lambda e: e["ts"])
    t0 = ev[0]["ts"]
    with open(path.replace(".json", ".txt"), "w") as f:
        for e in ev:
            f.write("%10.1f %9.1f %-13s %s\n" % (e["ts"] - t0, e["dur"], e["cat"], e["name"][:110]))
        gpu = [e for e in ev if e["cat"] in ("kernel", "gpu_memcpy", "gpu_memset")]
        busy = sum(e["dur"] for e in gpu)
        f.write("# GPU span %.1f us, busy %.1f us\n" % (gpu[-1]["ts"] + gpu[-1]["dur"] - gpu[0]["ts"], busy))
    os.remove(path)


run("nn", lambda: search(x))
run("finch", lambda: FINCH(x, backend=be, verbose=False, first_neighbors=search))

# event-timed repeats (no profiler): level-0 stage and full step, max over ranks
for label, fn in (("nn_stage", lambda: search(x)), ("finch", lambda: FINCH(x, backend=be, verbose=False, first_neighbors=search))):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 5], device=be.device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("%s: %.3f ms (max over %d ranks)" % (label, float(t), world), flush=True)
from video_similarity_search_b200.sharded import close_peer_groups
close_peer_groups()
dist.destroy_process_group()
