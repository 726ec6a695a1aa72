"""CUPTI timeline (torch.profiler) of one level-0 search and one full FINCH step: kernels, memcpys and CUDA runtime
calls with start / duration, written to gpurun_out/timeline_*.txt (diagnostic; never a bench number)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from video_similarity_search_b200 import synth
from video_similarity_search_b200.backend import CudaBackend
from video_similarity_search_b200.clustering.finch import FINCH

be = CudaBackend()
x = be.to_device(synth.config(sys.argv[1] if len(sys.argv) > 1 else "C3"))
for _ in range(3):
    FINCH(x, backend=be, verbose=False)
torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)


def run(name, fn):
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    path = "gpurun_out/timeline_%s.json" % name
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("ph") == "X" and e.get("cat") in
          ("kernel", "gpu_memcpy", "gpu_memset", "cuda_runtime", "cuda_driver")]
    ev.sort(key=lambda e: e["ts"])
    t0 = ev[0]["ts"]
    with open("gpurun_out/timeline_%s.txt" % name, "w") as f:
        for e in ev:
            f.write("%10.1f %9.1f %-13s %s\n" % (e["ts"] - t0, e["dur"], e["cat"], e["name"][:110]))
    os.remove(path)


run("nn", lambda: be.first_neighbors(x))
run("finch", lambda: FINCH(x, backend=be, verbose=False))
