"""bench.py contract pieces that can be checked without a GPU: the reference arm answers with one
JSON line (never a traceback) and the product arm refuses to run without CUDA instead of falling
back to a CPU path."""

from __future__ import annotations

import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
no_gpu = pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")


@no_gpu
def test_reference_arm_prints_one_json_line_without_a_gpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and "unavailable" in d


@no_gpu
def test_product_arm_has_no_cpu_fallback():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0
    assert "CUDA" in (r.stderr + r.stdout)


def test_bench_configs_cover_the_baseline_configs():
    sys.path.insert(0, ROOT)
    import bench

    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert {"cfg2", "cfg3", "cfg4", "cfg5"} <= set(bench.CONFIGS)
    assert bench.CONFIGS["cfg3"]["n_docs"] == 1_000_000 and bench.CONFIGS["cfg3"]["top_k"] == 100
    assert "north_star" in base
