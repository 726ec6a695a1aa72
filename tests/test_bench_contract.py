"""bench.py keeps the driver's JSON contract: the reference (CPU) arm is exercised here on a small workload; the B200
arm under -m gpu.  Sizes are chosen so that both finish in seconds - they check the line's shape, not its numbers."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(args, timeout=600):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_reference_arm_prints_the_contract_line():
    line = _run(["--impl", "reference", "--workload", "C1", "--steps", "1", "--warmup", "0", "--cpu-sample-rows", "512"])
    assert BASE_KEYS <= set(line) and line["impl"] == "reference"
    assert line["metric"] == "finch_full_hierarchy_embeddings_per_s" and line["unit"] == "embeddings/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["data"] == "synthetic"
    assert "workload" in line["config"] and "model" not in line["config"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    # (the timing arm feeds approximate level-0 neighbours as initial_rank - no min_sim cut - so only the first levels
    # coincide with the reference's [1170, 101, 25, 8, 5] for BASELINE configs[0])
    assert line["cpu_baseline"]["detail"]["num_clust"][:2] == [1170, 101]


@pytest.mark.gpu
def test_b200_arm_prints_the_contract_line():
    line = _run(["--workload", "C1", "--steps", "2", "--warmup", "3", "--cpu-sample-rows", "512"])
    assert BASE_KEYS <= set(line) and "impl" not in line
    assert {"gpu_launches", "clocks", "roofline", "nn_stage", "finch_seconds", "parity"} <= set(line)
    assert line["gpu_launches"] > 0 and line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] >= 3
    rf = line["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(rf) and rf["bound"] == "tensor"
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    e2e = line["e2e"]
    assert e2e["h2d_bytes_per_step"] == 9537 * 512 * 4 and e2e["d2h_bytes_per_step"] > 0 and e2e["value"] > 0
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(line["clocks"])
    assert line["config"]["partitions"] == [1170, 101, 25, 8, 5]
    assert line["parity"]["partition_equals_oracle"] is True
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["unit"] == line["unit"]
