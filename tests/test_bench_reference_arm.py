"""bench.py --impl reference runs without a GPU: it times the reference's CPU implementation of the
path (oracle/_ref when it was built from /root/reference, the plain-C port otherwise) and prints the
same JSON line as the GPU arm."""
import json
import os
import subprocess
import sys
import types

import pytest

from conftest import ROOT

KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e", "gpu_launches"}


def test_reference_arm_line():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_harness")):
        pytest.skip("oracle/_ref not built on this machine")
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                                   "--workload", "C1", "--cpu-scale", "0.05", "--steps", "1", "--warmup", "1",
                                   "--cpu-threads", "2"], text=True, timeout=300)
    line = json.loads(out.strip().splitlines()[-1])
    assert KEYS <= set(line), KEYS - set(line)
    assert line["impl"] == "reference" and line["metric"] == "nnz_per_s" and line["unit"] == "nnz/s"
    assert line["value"] > 0 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] == 2
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0


def test_port_arm_when_reference_is_not_built():
    sys.path.insert(0, ROOT)
    import bench
    args = types.SimpleNamespace(cpu_scale=0.05, workload="C1", seed=1, warmup=1, steps=1)
    shape, k, test_rows = bench.WORKLOADS["C1"]
    r = bench.run_port_arm(args, shape, k, test_rows)
    assert r["kind"] == "port" and r["cores"] == 1 and r["value"] > 0 and len(r["cg_iters"]) == 1
