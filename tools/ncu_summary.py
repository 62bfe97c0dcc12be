"""Summarise one kernel of an Nsight Compute report as a markdown table for profiles/.

    python tools/ncu_summary.py gpurun_out/foo.ncu-rep|foo.raw.csv --title "..." --command "..." \
        --workload "..." [--los N --evals N] > profiles/rN_ncu_foo.md

Reads the report with `ncu -i <rep> --page raw --csv` (no GPU needed) and keeps the metrics that
describe a compute-pipe bound kernel: duration, launch shape, occupancy, issue-slot / pipe
utilisations, instruction count, DRAM bytes and the largest stall reasons.
"""
from __future__ import annotations

import argparse
import csv
import io
import subprocess
import sys

KEEP = [
    "Kernel Name",
    "gpu__time_duration.sum",
    "launch__grid_size",
    "launch__block_size",
    "launch__registers_per_thread",
    "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "sm__cycles_elapsed.avg",
    "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "smsp__thread_inst_executed_pred_on_per_inst_executed.ratio",
]
STALL_PREFIX = "smsp__average_warps_issue_stalled_"
STALL_SUFFIX = "_per_issue_active.ratio"


def read_raw(report: str, kernel_index: int):
    if report.endswith(".csv"):  # already exported with `ncu -i <rep> --page raw --csv`
        text = open(report).read()
    else:
        text = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], check=True, capture_output=True,
                              text=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    header, units = rows[0], rows[1]
    data = rows[2 + kernel_index]
    return {h: (v, u) for h, u, v in zip(header, units, data)}


def to_bytes(value: str, unit: str) -> float:
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return float(value.replace(",", "")) * scale


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--kernel-index", type=int, default=0)
    ap.add_argument("--title", required=True)
    ap.add_argument("--command", required=True)
    ap.add_argument("--workload", required=True)
    ap.add_argument("--los", type=float, default=0.0, help="lines of sight in the launch")
    ap.add_argument("--evals", type=float, default=0.0, help="evaluations in the launch")
    ap.add_argument("--note", action="append", default=[])
    ap.add_argument("--counts-json", help="merge this kernel's executed counts per evaluation into this JSON "
                                          "file (read by bench.py; key = --counts-key)")
    ap.add_argument("--counts-key")
    ap.add_argument("--source", help="path of the markdown summary, recorded in the counts file")
    args = ap.parse_args()

    m = read_raw(args.report, args.kernel_index)
    out = [f"# {args.title}", "", f"Command (under gpurun, 1 GPU): `{args.command}`", "",
           f"Workload of this capture: {args.workload}", "", "| metric | value | unit |", "|---|---|---|"]
    for k in KEEP:
        if k in m:
            out.append(f"| {k} | {m[k][0]} | {m[k][1]} |")
    stalls = sorted(((float(v[0].replace(",", "")), k) for k, v in m.items()
                     if k.startswith(STALL_PREFIX) and k.endswith(STALL_SUFFIX) and v[0] not in ("", "n/a")),
                    reverse=True)[:8]
    for val, k in stalls:
        out.append(f"| {k} | {val:.6f} | {m[k][1]} |")
    out.append("")
    rd = to_bytes(*m["dram__bytes_read.sum"])
    wr = to_bytes(*m["dram__bytes_write.sum"])
    line = f"DRAM traffic of the launch: {rd / 1e6:.1f} MB read + {wr / 1e6:.1f} MB written = {(rd + wr) / 1e6:.1f} MB"
    if args.los:
        line += f" = {(rd + wr) / args.los:.1f} B per line of sight"
    out.append(line + ".")
    if args.evals:
        inst = float(m["smsp__inst_executed.sum"][0].replace(",", ""))
        out.append(f"Issue slots: {inst:.4g} warp-instructions for {args.evals:.4g} evaluations = "
                   f"{inst * 32 / args.evals:.1f} thread-instructions per evaluation.")
    out.extend(args.note)
    sys.stdout.write("\n".join(out) + "\n")
    if args.counts_json:
        write_counts(args, m, rd + wr)


def write_counts(args, m, dram_bytes) -> None:
    """Executed work per evaluation (units of bench.py): warp instructions, XU-pipe warp instructions,
    FMA / FP64 pipe cycles (SM sub-partition cycles), DRAM bytes per line of sight."""
    import json
    import os

    def num(key):
        return float(m[key][0].replace(",", "")) if key in m and m[key][0] not in ("", "n/a") else None

    if not (args.evals and args.los and args.counts_key):
        raise SystemExit("--counts-json needs --counts-key, --los and --evals")
    cycles, n_sm = num("sm__cycles_elapsed.avg"), None
    inst = num("smsp__inst_executed.sum")
    sub_cycles = None
    grid_sms = num("launch__sm_count") or num("device__attribute_multiprocessor_count")
    if cycles and grid_sms:
        sub_cycles = cycles * grid_sms * 4

    def pipe_cycles(pct_key):
        pct = num(pct_key)
        return None if pct is None or sub_cycles is None else pct / 100.0 * sub_cycles / args.evals

    xu_inst = num("sm__inst_executed_pipe_xu.sum")
    if xu_inst is None and sub_cycles is not None and num("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"):
        # one XU warp instruction occupies the pipe of an SM sub-partition for 8 cycles
        xu_inst = num("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active") / 100.0 * sub_cycles / 8.0
    entry = {
        "source": args.source, "kernel": m["Kernel Name"][0], "n_los": args.los, "evaluations": args.evals,
        "warp_inst_per_unit": inst / args.evals,
        "xu_warp_inst_per_unit": None if xu_inst is None else xu_inst / args.evals,
        "fma_pipe_cycles_per_unit": pipe_cycles("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        "fp64_pipe_cycles_per_unit": pipe_cycles("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "dram_bytes_per_los": dram_bytes / args.los,
    }
    data = {}
    if os.path.exists(args.counts_json):
        with open(args.counts_json) as fh:
            data = json.load(fh)
    data[args.counts_key] = entry
    with open(args.counts_json, "w") as fh:
        json.dump(data, fh, indent=1, sort_keys=True)
        fh.write("\n")


if __name__ == "__main__":
    main()
