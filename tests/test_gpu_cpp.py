"""C++ programs written against the header-only front end (the reference's C++ API shape), GPU."""
import json
import os
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def build(target):
    """Prebuilt by __graft_entry__.build() (`make tests`); only built here when missing, so that a
    loaded libgfb200.so is never relinked under a running test session."""
    exe = os.path.join(ROOT, "build", target)
    if not os.path.exists(exe):
        subprocess.run(["make", "build/" + target], cwd=ROOT, check=True, capture_output=True)
    return exe


def test_known_answers():
    """physics_test / solver_test / dispersion_test scenarios (tests/cpp/known_answers.cpp)."""
    exe = build("known_answers")
    out = subprocess.run([exe], cwd=ROOT, capture_output=True, text=True, timeout=1800)
    print(out.stdout[-3000:])
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert "0 failure(s)" in out.stdout
    assert out.stdout.count("ok   ") >= 20


def test_xrays_bench_program():
    """The reference benchmark's workload through the C++ API: kx solves to -500.0000036 and x moves inward."""
    exe = build("xrays_bench_b200")
    out = subprocess.run([exe, "20000", "200"], cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout + out.stderr
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    r = json.loads(line)
    assert r["rays"] == 20000 and r["ray_steps_per_s"] > 1.0e7
    ray0 = [l for l in out.stdout.splitlines() if l.startswith("ray 0")][0]
    assert "t = 1.000000" in ray0
