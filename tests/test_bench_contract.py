"""bench.py contract pieces that run without a GPU: the reference arm (the reference's own CPU
path through oracle/_ref) prints one well-formed JSON line, and the GPU arm refuses to run
without a device instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def test_reference_arm_prints_contract_line():
    from oracle import reference
    if not reference.available():
        pytest.skip("oracle/_ref not built")
    out = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-rays", "256", "--workload", "slab_omode"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "ray-steps/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_gpu_arm_has_no_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, "bench.py", "--steps", "1", "--warmup", "0"], cwd=ROOT,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode != 0
    assert "no CUDA device" in (out.stderr + out.stdout)
