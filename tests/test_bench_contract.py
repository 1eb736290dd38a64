"""bench.py's output contract, as far as it can be checked without a GPU: the reference arm prints exactly one JSON
line with the keys the driver reads, and the product arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-queens", "8"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "search_nodes_per_sec" and d["unit"] == "nodes/s"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["value"] > 0 and d["vs_baseline"] is None


def test_reference_arm_under_torchrun_only_rank0_prints():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0", "--cpu-queens", "6"], capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode != 0 and "no CUDA device" in (out.stderr + out.stdout)
    assert out.stdout.strip() == ""               # nothing that could be taken for a result


def test_roofline_traffic_comes_from_the_committed_captures():
    """`roofline.traffic` of the three workloads is read from profiles/r?_traffic.json (ncu --set full captures),
    keyed by the kernel the workload runs and the workload's name"""
    sys.path.insert(0, ROOT)
    import bench
    assert bench.measured_traffic("k_search_lov<false,true>", "queens16-all") == 7714048.0
    assert bench.measured_traffic("k_search_sat<false>", "sat200-seed1-any") > 1e6
    assert bench.measured_traffic("k_search<false,false,LIN=true>", "wcet-max") > 1e8
    assert bench.measured_traffic("k_search_lov<false,true>", "queens15-all") is None     # no capture of that workload
