#!/usr/bin/env python
"""bench.py - the hot path of BASELINE.json measured on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  (N > 1: launched by torch.distributed.run, one rank per GPU, NCCL)

Workload (config.workload): BASELINE configs[2]/[3] - FINCH full hierarchy on N = 240 000 x D = 512
synthetic Gaussian-mixture embeddings (Kinetics-400 train size), seed 0 (video_similarity_search_b200.synth).
A "step" is one complete FINCH call on that batch (all levels: first neighbours, components, means).

  value    embeddings/s through the whole hierarchy with the matrix already resident in HBM
           (= N / seconds per step; the BASELINE metric "FINCH full-hierarchy seconds" is `finch_seconds`)
  e2e      the same through the reference-facing call FINCH(numpy array): host->device copy of the
           embeddings and device->host copy of the label matrix inside the timed region
  roofline the dominant kernel (nn_screen_kernel, tcgen05): 2 * nq * n * d_pad flop per launch over its
           CUDA-event duration on its own stream, against MEASURED_PEAKS.json
  cpu_baseline the oracle port of the reference (numpy / scipy / sklearn) on this box's host cores, on a
           bounded sample (see `sample`)
At N > 1 the level-0 nearest-neighbour stage is row-sharded over the ranks (strong scaling: total work fixed).

--impl reference: the reference's CPU implementation (oracle port; the reference itself is pure Python that
needs pyflann above 70 000 rows, see oracle/finch_oracle.py) on the same workload, bounded sample per step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "C3"          # N=240000, D=512, K=400, seed 0
METRIC = "finch_full_hierarchy_embeddings_per_s"
UNIT = "embeddings/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=WORKLOAD, help="C1 | C3 | C5 (default C3; others for local experiments)")
    ap.add_argument("--cpu-sample-rows", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from the committed
# `ncu --set full` capture (profiles/r1_sym_screen_kernel_ncu_full.txt); keyed by (workload, ranks).  Not measured live:
# a number taken under the profiler's replay is evidence of traffic, never of time.
NCU_TRAFFIC_BYTES = {("C3", 1): 2.115234e9 + 102.021376e6}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(bf16_tflops=p["bf16_tflops"], bf16_tflops_sustained=p.get("bf16_tflops_sustained"),
                    hbm_gbs=p["hbm_gbs"], source="measured")
    return dict(bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, hbm_gbs=6650.0, source="fallback")


# --------------------------------------------------------------------------------------------------
# CPU side: the oracle port of the reference, bounded sample
# --------------------------------------------------------------------------------------------------
def within_component_neighbors(x, lab):
    """First neighbours restricted to rows of the same mixture component - a cheap way to obtain a realistic
    level-0 neighbour array for TIMING the CPU levels >= 1 (equal to the global first neighbour on all but a
    handful of rows; never used for a parity claim)."""
    from oracle import finch_oracle as fo
    nn = np.empty(len(x), dtype=np.int64)
    order = np.argsort(lab, kind="stable")
    bounds = np.flatnonzero(np.diff(lab[order])) + 1
    for grp in np.split(order, bounds):
        if len(grp) == 1:
            nn[grp] = grp
            continue
        j, _, _ = fo.first_neighbors_blocked(x[grp])
        nn[grp] = grp[j]
    return nn


def cpu_reference_step(x, nn_for_rest, sample_rows):
    """One bounded CPU step of the reference algorithm on the workload:
      (a) level-0 first-neighbour stage (finch.py:22-29 arithmetic, exact blocked form used above 70 000 rows)
          on `sample_rows` query rows x ALL columns, scaled linearly to all rows;
      (b) everything after it - components, means, all levels >= 1 - in full (oracle finch with initial_rank).
    Returns (estimated seconds for the whole hierarchy, detail dict)."""
    from oracle import finch_oracle as fo
    n = len(x)
    rows = np.linspace(0, n - 1, sample_rows).astype(np.int64)
    t0 = time.perf_counter()
    fo.first_neighbors_blocked(x, rows=rows)
    t_nn = time.perf_counter() - t0
    t0 = time.perf_counter()
    c, num_clust, _ = fo.finch(x, initial_rank=nn_for_rest)
    t_rest = time.perf_counter() - t0
    est = t_nn * (n / float(sample_rows)) + t_rest
    return est, dict(nn_sample_s=t_nn, nn_stage_est_s=t_nn * n / float(sample_rows), rest_s=t_rest,
                     num_clust=[int(v) for v in num_clust])


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from video_similarity_search_b200 import synth
    n, d, k, seed = synth.CONFIGS[args.workload]
    x, lab, _ = synth.gaussian_mixture(n, d, k, seed, return_labels=True)
    nn = within_component_neighbors(x, lab)
    sample = min(args.cpu_sample_rows, n)
    times, detail = [], None
    for i in range(args.warmup + args.steps):
        est, detail = cpu_reference_step(x, nn, sample)
        if i >= args.warmup:
            times.append(est)
    sec = statistics.mean(times)
    cores = blas_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": n / sec, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "FINCH full hierarchy, N=%d x D=%d Gaussian mixture (K=%d, seed %d), cosine" % (n, d, k, seed)},
        "finch_seconds": sec,
        "cpu_baseline": {"value": n / sec, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "level-0 NN stage on %d of %d query rows x all columns, scaled linearly in rows; "
                                   "levels >= 1, components and means timed in full (oracle port of finch.py with "
                                   "the exact-NN stand-in the reference needs above 70 000 rows)" % (sample, n),
                         "detail": detail},
        "e2e": {"value": n / sec, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------------
# GPU side
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled from a thread every ~10 ms
    (nvidia-smi -lms 100 as the fallback when pynvml cannot initialise)."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.thread = None
        self.stop_flag = False
        self.sm, self.power, self.reasons, self.max_mhz = [], [], set(), None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.gpu_index < len(ids) and ids[self.gpu_index].isdigit():
                return int(ids[self.gpu_index])
        return self.gpu_index

    def _poll(self, nv, handle):
        names = [("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown),
                 ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown),
                 ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap),
                 ("hw_power_brake_slowdown", nv.nvmlClocksEventReasonHwPowerBrakeSlowdown)]
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(handle) / 1000.0)
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(handle)
                for name, bit in names:
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            handle = nv.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, args=(nv, handle), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self._physical_index()), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                    "samples": len(self.sm), "power_w_max": max(self.power) if self.power else None,
                    "reasons": sorted(self.reasons), "how": "NVML polled every 10 ms during the timed steps"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "how": "nvidia-smi -lms 100 during the timed steps"}


def run_b200(args):
    import torch
    import torch.distributed as dist
    from video_similarity_search_b200 import _lib, synth
    from video_similarity_search_b200.backend import CudaBackend
    from video_similarity_search_b200.clustering.finch import FINCH
    from video_similarity_search_b200.sharded import FINCH_sharded, sharded_first_neighbors

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    be = CudaBackend()
    lib = _lib.load()

    n, d, k, seed = synth.CONFIGS[args.workload]
    x_host, lab_host, _ = synth.gaussian_mixture(n, d, k, seed, return_labels=True)
    x_pinned = torch.from_numpy(x_host).pin_memory()
    x_dev = x_pinned.to(be.device, non_blocking=True)
    torch.cuda.synchronize()
    search = sharded_first_neighbors(be) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=be.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps):
        """EXACTLY `steps` calls between barrier+synchronize, CUDA events on the current stream, max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / steps, out

    def step_resident():
        return FINCH(x_dev, verbose=False, backend=be, first_neighbors=search)

    def step_e2e():
        # the call a reference user makes: host matrix in, numpy out (H2D of the embeddings, D2H of the labels inside).
        # N > 1: FINCH_sharded - every rank uploads 1 / N of the rows and an all-gather over NVLink assembles the matrix
        if world > 1:
            return FINCH_sharded(x_pinned, verbose=False, backend=be)
        return FINCH(x_pinned, verbose=False, backend=be)

    def step_nn_only():
        return (search(x_dev) if search else be.first_neighbors(x_dev))

    # warm-up (also builds the stream-ordered memory pool)
    for _ in range(max(args.warmup, 3)):
        step_resident()
    lib.slic_profile_screen(1)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lib.slic_launch_count()
    ms_step, result = timed(step_resident, args.steps)
    launches = (lib.slic_launch_count() - launches0)
    clocks = sampler.stop() if rank == 0 else None
    c, num_clust, _ = result

    # dominant kernel: per-launch CUDA-event time of nn_screen_kernel at level 0 (the last profiled launches are the
    # small levels, so time the level-0 search alone, K launches, events recorded around the kernel on its stream)
    import ctypes
    screen_ms = []
    flop, exec_flop = ctypes.c_double(0), ctypes.c_double(0)
    for _ in range(args.steps):
        step_nn_only()
        # the level-0 call is the only screen launch in step_nn_only
        ms = ctypes.c_float(0)
        _lib.check(lib.slic_last_screen_time(ctypes.byref(ms), ctypes.byref(flop)), "slic_last_screen_time")
        _lib.check(lib.slic_last_screen_exec_flop(ctypes.byref(exec_flop)), "slic_last_screen_exec_flop")
        screen_ms.append(ms.value)
    lib.slic_profile_screen(0)
    screen_ms_avg = max_over_ranks(statistics.mean(screen_ms))
    ms_nn, _ = timed(step_nn_only, args.steps)
    for _ in range(2):
        step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)

    sharded_equal = None
    if world > 1:
        # parity of the exchange step, outside the timed region: the merged multi-GPU first neighbours must equal the
        # single-GPU search of the same matrix on this rank, bit for bit (indices and float32 distances)
        nn_s, d_s, _ = search(x_dev)
        nn_1, d_1, _ = be.first_neighbors(x_dev)
        ok = torch.tensor([int(torch.equal(nn_s, nn_1) and torch.equal(d_s, d_1))], device=be.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        sharded_equal = bool(ok.item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = load_peaks()
    achieved_tf = flop.value / (screen_ms_avg * 1e-3) / 1e12
    executed_tf = exec_flop.value / (screen_ms_avg * 1e-3) / 1e12
    d_pad = (d + 63) // 64 * 64
    line = {
        "metric": METRIC, "value": n / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16 screen + f32/f64 exact re-rank", "data": "synthetic",
        "config": {"workload": "FINCH full hierarchy, N=%d x D=%d Gaussian mixture (K=%d, seed %d), cosine" % (n, d, k, seed),
                   "l2": "inputs exceed L2 (%.0f MB fp32 + %.0f MB bf16 per step)" % (n * d * 4 / 1e6, n * d_pad * 2 / 1e6),
                   "partitions": [int(v) for v in num_clust],
                   "parallelism": ("1 GPU" if world == 1 else
                                   "level-0 NN: the %d ranks share the tiles of the symmetric screen's triangle, (distance, "
                                   "neighbour) keys merged by one NCCL all-reduce MIN of 8(N+1) bytes; levels >= 1 replicated"
                                   % world)},
        "finch_seconds": ms_step * 1e-3,
        "nn_stage": {"ms": ms_nn, "queries_per_s": n / (ms_nn * 1e-3)},
        "e2e": {"value": n / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                # whole job: every rank uploads its 1 / N share of the rows (the rest arrives over NVLink); labels back per rank
                "h2d_bytes_per_step": int(n * d * 4), "d2h_bytes_per_step": int(c.size * 4 * world)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"kernel": "nn_screen_kernel (level 0, %d x %d x %d%s)" % (n, n, d_pad, "" if world == 1 else
                                                                               ", this rank's 1/%d of the tiles" % world),
                     "bound": "tensor", "achieved": achieved_tf, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                     "frac": achieved_tf / peaks["bf16_tflops"], "peak_source": peaks["source"] + " burst bf16",
                     "frac_of_sustained": (achieved_tf / peaks["bf16_tflops_sustained"]) if peaks.get("bf16_tflops_sustained") else None,
                     "kernel_ms": screen_ms_avg, "flop_per_launch": flop.value,
                     # the self-search computes only the tiles on or right of the diagonal of the symmetric score matrix
                     # (each filtered along rows AND columns): `achieved` counts the ALGORITHMIC 2 n^2 d flop of SURVEY.md
                     # section 8(d) (full square, no symmetry discount) and may therefore exceed the peak; `executed` is
                     # what the tensor cores actually did, and `frac_executed` its fraction of the measured peak
                     "executed_flop_per_launch": exec_flop.value, "executed": executed_tf,
                     "frac_executed": executed_tf / peaks["bf16_tflops"],
                     "traffic": NCU_TRAFFIC_BYTES.get((args.workload, world)), "traffic_unit": "bytes/launch (ncu dram read+write)"},
    }
    if world == 1 and not args.no_cpu_baseline:
        nn_dev, _, _ = be.first_neighbors(x_dev)
        sample = min(args.cpu_sample_rows, n)
        # parity gate in the same run: oracle first neighbours on the sampled rows must equal the GPU's
        from oracle import finch_oracle as fo
        rows = np.linspace(0, n - 1, sample).astype(np.int64)
        t0 = time.perf_counter()
        enn, _, gap = fo.first_neighbors_blocked(x_host, rows=rows)
        t_nn = time.perf_counter() - t0
        nn_host = nn_dev.cpu().numpy().astype(np.int64)
        clear = gap > 2e-6
        parity_nn = bool(np.array_equal(nn_host[rows][clear], enn[clear]))
        t0 = time.perf_counter()
        if n <= fo.FLANN_THRESHOLD:
            # dense mode of the reference (distances kept, min_sim cut): the oracle runs it in full, only its level-0
            # neighbours are replaced by the GPU's so that float32 tie rows cannot change the partition
            co, no, _ = fo.finch(x_host, nn0_override=nn_host)
        else:
            co, no, _ = fo.finch(x_host, initial_rank=nn_host)
        t_rest = time.perf_counter() - t0
        parity_partition = bool(no == num_clust and np.array_equal(co, c))
        est = t_nn * n / float(sample) + t_rest
        line["cpu_baseline"] = {
            "value": n / est, "unit": UNIT, "cores": blas_threads(), "kind": "port",
            "sample": "level-0 NN stage on %d of %d query rows x all columns (%.1f s), scaled linearly in rows; levels >= 1, "
                      "components and means timed in full (%.1f s) - oracle port of finch.py with the exact-NN stand-in the "
                      "reference needs above 70 000 rows" % (sample, n, t_nn, t_rest),
            "finch_seconds_est": est}
        line["parity"] = {"first_neighbors_equal_on_sample": parity_nn, "tie_rows_in_sample": int((~clear).sum()),
                          "partition_equals_oracle": parity_partition}
    if sharded_equal is not None:
        line["parity"] = {"sharded_first_neighbors_equal_single_gpu_on_every_rank": sharded_equal}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
