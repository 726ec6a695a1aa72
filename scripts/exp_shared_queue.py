"""A/B of the multi-GPU level-0 search inside ONE process group (diagnostic): fixed 1/G shares against the box-wide unit
queue (SLIC_COMM_SHARED_QUEUE), unit lengths of the triangle (SLIC_SYM_UNIT_TILES) and of the fused pre-pass
(SLIC_SYM_PRE_TILES).  The library reads these variables per call, so one start-up serves every variant.
Launch: python -m torch.distributed.run --nproc-per-node G scripts/exp_shared_queue.py [C3] [steps]
Prints, per variant: level-0 stage and full hierarchy (CUDA events, max over ranks), screen kernel per rank, and whether
the merged neighbours equal the first variant's on this rank."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from video_similarity_search_b200 import _lib, synth
from video_similarity_search_b200.backend import CudaBackend
from video_similarity_search_b200.clustering.finch import FINCH
from video_similarity_search_b200.sharded import sharded_first_neighbors, close_peer_groups

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
be = CudaBackend()
lib = _lib.load()
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
x = be.to_device(synth.config(name))
search = sharded_first_neighbors(be)

VARIANTS = [
    ("fixed shares", {"SLIC_COMM_SHARED_QUEUE": "0"}),
    ("shared queue", {"SLIC_COMM_SHARED_QUEUE": "1"}),
    ("shared queue, units 32", {"SLIC_COMM_SHARED_QUEUE": "1", "SLIC_SYM_UNIT_TILES": "32"}),
    ("shared queue, units 16", {"SLIC_COMM_SHARED_QUEUE": "1", "SLIC_SYM_UNIT_TILES": "16"}),
    ("shared queue, units 32, pre-pass 8", {"SLIC_COMM_SHARED_QUEUE": "1", "SLIC_SYM_UNIT_TILES": "32", "SLIC_SYM_PRE_TILES": "8"}),
    ("fixed shares, units 32", {"SLIC_COMM_SHARED_QUEUE": "0", "SLIC_SYM_UNIT_TILES": "32"}),
    ("fixed shares (again)", {"SLIC_COMM_SHARED_QUEUE": "0"}),
    ("shared queue (again)", {"SLIC_COMM_SHARED_QUEUE": "1"}),
]
KEYS = ("SLIC_COMM_SHARED_QUEUE", "SLIC_SYM_UNIT_TILES", "SLIC_SYM_PRE_TILES")


def timed(fn, k):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / k], device=be.device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def nn_dropped():
    search(x)
    return None


for _ in range(15):   # start-up settling (see bench.py)
    FINCH(x, backend=be, verbose=False, first_neighbors=search)
ref = None
for label, env in VARIANTS:
    for k in KEYS:
        os.environ.pop(k, None)
    os.environ.update(env)
    for _ in range(3):
        nn_dropped()
    nn, _, _ = search(x)
    nn = nn.clone()
    if ref is None:
        ref = nn
    same = torch.tensor([1 if torch.equal(nn, ref) else 0], device=be.device)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    ms_nn = timed(nn_dropped, steps)
    ms_finch = timed(lambda: FINCH(x, backend=be, verbose=False, first_neighbors=search), steps)
    lib.slic_profile_screen(1)
    scr = []
    for _ in range(steps):
        nn_dropped()
        ms = ctypes.c_float(0); flop = ctypes.c_double(0)
        lib.slic_last_screen_time(ctypes.byref(ms), ctypes.byref(flop))
        scr.append(ms.value)
    lib.slic_profile_screen(0)
    t = torch.tensor([sum(scr) / len(scr)], device=be.device, dtype=torch.float64)
    tmax, tmin = t.clone(), t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("%-38s level-0 stage %.3f ms | hierarchy %.3f ms | screen kernel per rank %.3f-%.3f ms | neighbours equal: %s"
              % (label, ms_nn, ms_finch, float(tmin), float(tmax), bool(int(same))), flush=True)
close_peer_groups()
dist.destroy_process_group()
