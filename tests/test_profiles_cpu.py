"""The roofline bookkeeping that runs without a GPU: the committed ncu captures bench.py reads are
self-consistent, the converter that writes them parses ncu's raw page, and bench.py's selection /
staleness logic does what DESIGN.md section 4 says."""
import csv
import glob
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

WORKLOADS = ("efit_xmode", "efit_cold", "efit_absorb", "vmec_omode", "boris")


@pytest.mark.parametrize("workload", WORKLOADS)
def test_committed_capture_is_self_consistent(workload):
    with open(os.path.join(ROOT, "profiles", "r2_ncu_%s.json" % workload)) as f:
        cap = json.load(f)
    for key in ("kernel", "duration_ms", "dram_bytes_per_launch", "fp64_pipe_active_pct", "thread_inst", "units_per_launch",
                "unit_steps", "fp64_flop_per_unit_step", "kernel_text_sha256", "git_sha", "plain_run_ms_per_step", "stalls_per_issue"):
        assert cap.get(key) is not None, key
    t = cap["thread_inst"]
    unit_steps = cap["units_per_launch"]*cap["unit_steps"]
    assert abs(cap["fp64_flop_per_unit_step"] - (2.0*t["dfma"] + t["dadd"] + t["dmul"])/unit_steps) < 1.0e-6
    # executed flops / duration can never exceed the nominal FP64 pipe peak (148 SMs x 64 lanes x 2 x 1.965 GHz)
    tflops = (2.0*t["dfma"] + t["dadd"] + t["dmul"])/(cap["duration_ms"]*1.0e-3)/1.0e12
    assert 10.0 < tflops < 37.3, tflops
    # the capture ran under ncu (cold caches, serialised): within 5 % of the plain run of the same command
    assert abs(cap["duration_ms"]/cap["plain_run_ms_per_step"] - 1.0) < 0.05
    assert len(cap["kernel_text_sha256"]) == 64
    assert 40.0 < cap["fp64_pipe_active_pct"] < 100.0


def test_fp64_probes_bracket_the_step_kernels():
    """Nominal 37.2 > uniform-operand DFMA probe (92 % pipe active) > step kernels (76-80 %) > register-operand probe (67 %)."""
    caps = {}
    for name in ("fp64_peak", "fp64_peak_registers", "efit_xmode", "efit_cold"):
        with open(os.path.join(ROOT, "profiles", "r2_ncu_%s.json" % name)) as f:
            caps[name] = json.load(f)["fp64_pipe_active_pct"]
    assert caps["fp64_peak"] > caps["efit_cold"] > caps["fp64_peak_registers"]
    assert caps["fp64_peak"] > caps["efit_xmode"] > caps["fp64_peak_registers"]
    assert caps["fp64_peak"] > 90.0 and caps["fp64_peak_registers"] < 70.0


def test_ncu_to_json_parses_a_raw_page(tmp_path):
    """tools/ncu_to_json.py on a hand-made raw page: unit conversion, per-cycle rates x cycles = counts."""
    hdr = ["ID", "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__cycles_elapsed.avg",
           "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
           "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
           "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed", "launch__registers_per_thread"]
    units = ["", "", "ms", "Mbyte", "Kbyte", "%", "cycle", "inst/cycle", "inst/cycle", "inst/cycle", "register/thread"]
    rows = [["0", "other_kernel", "1.0", "1", "1", "10", "100", "1", "1", "1", "32"],
            ["1", "solver_kernel", "2.5", "64", "500", "77.5", "1000000", "600", "50", "300", "128"]]
    raw = tmp_path / "raw.csv"
    with open(raw, "w", newline="") as f:
        w = csv.writer(f, quoting=csv.QUOTE_ALL)
        w.writerow(hdr)
        w.writerow(units)
        w.writerows(rows)
    out = tmp_path / "cap.json"
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_to_json.py"), str(raw), "--workload", "w", "--kernel", "solver_kernel",
                    "--units", "1000", "--unit-steps", "100", "--out", str(out)], check=True, cwd=ROOT, capture_output=True)
    cap = json.load(open(out))
    assert cap["duration_ms"] == 2.5 and cap["dram_bytes_read"] == 64.0e6 and cap["dram_bytes_write"] == 500.0e3
    assert cap["dram_bytes_per_launch"] == 64.5e6 and cap["registers_per_thread"] == 128.0
    assert cap["thread_inst"] == {"dfma": 6.0e8, "dadd": 5.0e7, "dmul": 3.0e8}
    assert abs(cap["fp64_flop_per_unit_step"] - (2*6.0e8 + 5.0e7 + 3.0e8)/1.0e5) < 1.0e-9
    assert abs(cap["fma_share_of_fp64_arithmetic"] - 6.0/9.5) < 1.0e-12


def test_bench_picks_the_newest_capture_and_flags_staleness(tmp_path, monkeypatch):
    sys.path.insert(0, ROOT)
    import bench
    cap = bench.ncu_capture("efit_xmode", "0"*64)
    assert cap is not None and cap["file"].startswith("profiles/r2_ncu_efit_xmode") and cap["stale"] is True
    with open(os.path.join(ROOT, cap["file"])) as f:
        sha = json.load(f)["kernel_text_sha256"]
    assert bench.ncu_capture("efit_xmode", sha)["stale"] is False
    assert bench.ncu_capture("no_such_workload", sha) is None
    # every workload of the default line has a capture
    for w in WORKLOADS:
        assert glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_%s.json" % w)), w
