#!/usr/bin/env python3
"""Turn the raw page of an `ncu --set full` capture into the small JSON bench.py reads
(profiles/r<round>_ncu_<workload>.json): per-launch time, DRAM bytes, FP64 pipe share, the executed
FP64 instruction mix per unit-step, stalls.  Records the git sha and the sha256 of the kernel text the
capture was taken from, so that bench.py can say whether a capture still describes the binary it times.

    ncu -i prof.ncu-rep --page raw --csv > raw.csv
    tools/ncu_to_json.py raw.csv --workload efit_xmode --kernel solver_kernel --units 1000000 --unit-steps 100 \
        --bench-line plain.json --out profiles/r2_ncu_efit_xmode.json
"""
import argparse
import csv
import json
import subprocess
import sys


def number(text):
    try:
        return float(text.replace(",", ""))
    except ValueError:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw")
    ap.add_argument("--workload", required=True)
    ap.add_argument("--kernel", required=True, help="substring of the kernel name")
    ap.add_argument("--units", type=float, required=True, help="rays / particles one launch advances")
    ap.add_argument("--unit-steps", type=float, required=True, help="steps per launch")
    ap.add_argument("--bench-line", default=None, help="JSON line of the plain run of the same command (kernel text sha)")
    ap.add_argument("--command", default="")
    ap.add_argument("--out", required=True)
    args = ap.parse_args()

    rows = list(csv.reader(open(args.raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    matching = [r for r in rows[2:] if args.kernel in r[idx["Kernel Name"]]]
    if not matching:
        sys.exit("no launch of %s in %s" % (args.kernel, args.raw))
    r = matching[-1]

    def get(name):
        return number(r[idx[name]]) if name in idx else None

    def scaled(name, want):
        """value converted to `want` units (ncu picks ms/us, Mbyte/Gbyte ... per row)."""
        v, u = get(name), units[idx[name]] if name in idx else ""
        if v is None:
            return None
        scale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9,
                 "cycle": 1.0, "inst": 1.0, "": 1.0}.get(u, 1.0)
        return v*scale/want

    cycles = get("smsp__cycles_elapsed.avg") or get("sm__cycles_elapsed.avg")

    def thread_inst(op):
        """Thread-level executed instructions of one opcode class over the launch: ncu's --set full
        carries the per-elapsed-cycle rate summed over the GPU; x elapsed cycles = the count."""
        direct = get("smsp__sass_thread_inst_executed_op_%s_pred_on.sum" % op)
        if direct is not None:
            return direct
        rate = get("smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % op)
        return rate*cycles if rate is not None and cycles else None

    dfma, dadd, dmul = thread_inst("dfma"), thread_inst("dadd"), thread_inst("dmul")
    unit_steps = args.units*args.unit_steps
    warp_fp64 = get("sm__inst_executed_pipe_fp64.sum")
    out = {
        "workload": args.workload, "kernel": r[idx["Kernel Name"]], "command": args.command,
        "units_per_launch": args.units, "unit_steps": args.unit_steps,
        "duration_ms": scaled("gpu__time_duration.sum", 1e-3),
        "dram_bytes_read": scaled("dram__bytes_read.sum", 1.0), "dram_bytes_write": scaled("dram__bytes_write.sum", 1.0),
        "fp64_pipe_active_pct": get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "fp64_pipe_active_pct_of_elapsed": get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        "issue_active_pct": get("smsp__issue_active.avg.pct"),
        "warps_active_pct": get("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "registers_per_thread": get("launch__registers_per_thread"),
        "block_size": get("launch__block_size"), "grid_size": get("launch__grid_size"),
        "local_load_inst": get("smsp__sass_inst_executed_op_local_ld.sum"), "local_store_inst": get("smsp__sass_inst_executed_op_local_st.sum"),
        "l1_hit_pct": get("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": get("lts__t_sector_hit_rate.pct"),
        "warp_inst_executed": get("smsp__inst_executed.sum"),
        "thread_inst": {"dfma": dfma, "dadd": dadd, "dmul": dmul},
        "stalls_per_issue": {k: get("smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % k)
                             for k in ("wait", "math_pipe_throttle", "long_scoreboard", "short_scoreboard", "not_selected",
                                       "dispatch_stall", "no_instruction", "branch_resolving", "barrier", "lg_throttle", "mio_throttle")},
        "sm_clock_ghz": get("sm__cycles_elapsed.avg.per_second"),
    }
    if out["dram_bytes_read"] is not None and out["dram_bytes_write"] is not None:
        out["dram_bytes_per_launch"] = out["dram_bytes_read"] + out["dram_bytes_write"]
    if None not in (dfma, dadd, dmul):
        out["fp64_flop_per_unit_step"] = (2.0*dfma + dadd + dmul)/unit_steps
        out["fp64_thread_inst_per_unit_step"] = (dfma + dadd + dmul)/unit_steps
        out["fma_share_of_fp64_arithmetic"] = dfma/(dfma + dadd + dmul)
    if warp_fp64 is not None:
        out["fp64_inst_per_unit_step"] = warp_fp64*32.0/unit_steps       # every FP64-pipe instruction incl. compares, conversions, MUFU.*64H feeds
    if out["warp_inst_executed"]:
        out["inst_per_unit_step"] = out["warp_inst_executed"]*32.0/unit_steps
    if args.bench_line:
        with open(args.bench_line) as f:
            line = json.loads(f.read().strip().splitlines()[-1])
        out["kernel_text_sha256"] = line.get("roofline", {}).get("kernel_text_sha256")
        out["plain_run_ms_per_step"] = line.get("ms_per_step")
        out["plain_run_value"] = line.get("value")
    try:
        out["git_sha"] = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip() or None
    except OSError:
        out["git_sha"] = None
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
        f.write("\n")
    print("wrote", args.out, {k: out.get(k) for k in ("duration_ms", "fp64_pipe_active_pct", "fp64_flop_per_unit_step", "dram_bytes_per_launch")})


if __name__ == "__main__":
    main()
