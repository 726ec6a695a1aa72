"""Measurement of the section-8f kernels at Kinetics size (N = 240 000): device time by CUDA events against the
algorithmic bytes of DESIGN.md, with the CPU path (scikit-learn / the reference's Python loop) timed beside it.
Prints one JSON line; diagnostic, not the bench.py contract."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import metrics_oracle as mo
from video_similarity_search_b200 import metrics, synth
from video_similarity_search_b200.backend import CudaBackend
from video_similarity_search_b200.clustering.finch import FINCH

be = CudaBackend()
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}
x, lab, _ = synth.gaussian_mixture(240000, 512, 400, 0, return_labels=True)
xd = be.to_device(x)
c, num, _ = FINCH(xd, backend=be, verbose=False)
pred = c[:, 0]


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


out = {"n": 240000, "clusters": int(num[0]), "hbm_peak_gbs": peaks["hbm_gbs"]}
lt = be.to_device(lab.astype(np.int32), torch.int32)
lp = be.to_device(pred.astype(np.int32), torch.int32)
R, C = 400, int(num[0])
ms, _ = timed(lambda: be.cluster_metrics(lt, lp, R, C, True))
t0 = time.perf_counter(); ami_cpu = mo.adjusted_mutual_info_score(lab, pred); t_ami = time.perf_counter() - t0
t0 = time.perf_counter(); nmi_cpu = mo.normalized_mutual_info_score(lab, pred); t_nmi = time.perf_counter() - t0
out["cluster_metrics"] = {"ms": ms, "algorithmic_bytes": 8 * 240000 + 4 * R * C, "cpu_sklearn_ami_s": t_ami, "cpu_sklearn_nmi_s": t_nmi,
                          "ami": metrics.adjusted_mutual_info_score(lab, pred, backend=be), "ami_cpu": ami_cpu,
                          "nmi": metrics.normalized_mutual_info_score(lab, pred, backend=be), "nmi_cpu": nmi_cpu}
ms, _ = timed(lambda: be.cluster_metrics(lt, lp, R, C, False))
out["cluster_metrics"]["ms_without_emi"] = ms
from video_similarity_search_b200 import _lib
cen = torch.empty_like(xd)
ms, _ = timed(lambda: _lib.call("slic_center_columns", xd.data_ptr(), 240000, 512, cen.data_ptr(), None, None))
out["center_columns"] = {"ms": ms, "algorithmic_bytes": 3 * x.nbytes, "gbs": 3 * x.nbytes / ms / 1e6,
                         "frac_of_hbm_peak": 3 * x.nbytes / ms / 1e6 / peaks["hbm_gbs"]}
idx = torch.randperm(240000, device=be.device)
ms, _ = timed(lambda: be.scatter_last_wins(lp, idx, 240000))
t0 = time.perf_counter(); mo.unshuffled_assignments(pred.tolist(), idx.cpu().tolist(), 240000); t_py = time.perf_counter() - t0
out["scatter_last_wins"] = {"ms": ms, "algorithmic_bytes": 12 * 240000 + 8 * 240000, "cpu_python_loop_s": t_py}
unit_buf, bf_buf = torch.empty_like(xd), torch.empty((240000, 512), dtype=torch.bfloat16, device=xd.device)
ms, _ = timed(lambda: _lib.call("slic_normalize_rows", xd.data_ptr(), 240000, 512, 0, unit_buf.data_ptr(), None,
                                bf_buf.data_ptr(), 512, None))
out["normalize_rows"] = {"ms": ms, "algorithmic_bytes": int(x.nbytes * 2.5), "gbs": x.nbytes * 2.5 / ms / 1e6,
                         "frac_of_hbm_peak": x.nbytes * 2.5 / ms / 1e6 / peaks["hbm_gbs"]}
ms, _ = timed(lambda: be.cluster_sums(xd, lp, C))
out["cluster_sums_level0"] = {"ms": ms, "algorithmic_bytes": x.nbytes + 4 * 240000 + 16 * C * 512,
                              "gbs": (x.nbytes + 4 * 240000 + 16 * C * 512) / ms / 1e6,
                              "frac_of_hbm_peak": (x.nbytes + 4 * 240000 + 16 * C * 512) / ms / 1e6 / peaks["hbm_gbs"]}
print(json.dumps(out))
