#!/usr/bin/env python
"""Turn an ncu report (+ optional launch-list CSV) into the markdown summary kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--launches gpurun_out/launches.csv]
                                [--kernel-index N] [--title "..."] > profiles/<name>.md

Runs `ncu -i` (no GPU needed) for the raw page of every captured launch and the SASS+CUDA source
page of one launch (default: the longest), and prints: per-launch headline metrics, stall reasons,
samples / instructions per source file and the hottest source lines.
"""
import argparse
import csv
import io
import subprocess
import sys
from collections import defaultdict

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__grid_size", "grid"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.per_cycle_active", "issue slots busy (IPC per SMSP)"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle / SMSP"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads per warp instruction"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe % of peak"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe % of peak"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe % of peak"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe % of peak"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % (ncu SOL)"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % (ncu SOL)"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
]


def ncu(*args):
    return subprocess.run(["ncu", *args], check=True, capture_output=True, text=True).stdout


def fnum(x):
    try:
        return float(str(x).replace(",", ""))
    except ValueError:
        return 0.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--launches")
    ap.add_argument("--kernel-index", type=int, default=None, help="1-based index of the launch for the source page")
    ap.add_argument("--kernel-regex", default="sp_level")
    ap.add_argument("--title", default="ncu summary")
    ap.add_argument("--top", type=int, default=40)
    args = ap.parse_args()

    print(f"# {args.title}\n")
    print(f"Source: `{args.report}` (`ncu --set full --clock-control none --import-source on`), read offline with "
          "`ncu -i ... --page raw|source --csv`.\n")

    if args.launches:
        rows = [r for r in csv.reader(open(args.launches)) if len(r) > 5]
        hdr = rows[0]
        ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
        mi = hdr.index("Metric Name") if "Metric Name" in hdr else None
        tot, cnt = defaultdict(float), defaultdict(int)
        for r in rows[1:]:
            if mi is not None and r[mi] != "gpu__time_duration.sum":
                continue                                    # launch lists that carry several metrics per launch
            name = r[ki].split("(")[0]
            if name.startswith(("sp_ffma", "sp_copy")):
                continue                                    # roofline micro-benchmarks of bench.py, not part of a frame
            tot[name] += fnum(r[vi]); cnt[name] += 1
        s = sum(tot.values())
        print(f"## Launch list (`{args.launches}`, gpu__time_duration.sum; cold-cache, serialised: compare shares)\n")
        print("| kernel | launches | total ms | share |\n|---|---|---|---|")
        for k in sorted(tot, key=lambda k: -tot[k]):
            print(f"| {k} | {cnt[k]} | {tot[k] / 1e6:.3f} | {tot[k] / s:.1%} |")
        print()

    raw = list(csv.reader(io.StringIO(ncu("-i", args.report, "--page", "raw", "--csv"))))
    hdr, units, data = raw[0], raw[1], raw[2:]
    print("## Per-launch metrics (raw page)\n")
    print("| metric | unit | " + " | ".join(f"launch {i + 1}" for i in range(len(data))) + " |")
    print("|---|---|" + "---|" * len(data))
    for key, label in METRICS:
        if key in hdr:
            i = hdr.index(key)
            print(f"| {label} (`{key}`) | {units[i]} | " + " | ".join(r[i] for r in data) + " |")
    dur_i = hdr.index("gpu__time_duration.sum")
    pick = args.kernel_index or (max(range(len(data)), key=lambda j: fnum(data[j][dur_i])) + 1)
    row = data[pick - 1]
    print(f"\n## Warp stall reasons, launch {pick} (warps stalled per issue-active cycle)\n")
    stalls = [(h, fnum(row[i])) for i, h in enumerate(hdr)
              if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    print("| reason | ratio |\n|---|---|")
    for h, v in sorted(stalls, key=lambda x: -x[1])[:10]:
        print(f"| {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} | {v:.3f} |")

    src = ncu("-i", args.report, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id",
              f"::regex:{args.kernel_regex}:{pick}")
    rows, cur, h2 = [], None, None
    for r in csv.reader(io.StringIO(src)):
        if r and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            h2 = r
        elif h2 and len(r) == len(h2) and r[2] == "-":
            rows.append((cur, r))
    if rows:
        iS, iI, iT = h2.index("# Samples"), h2.index("Instructions Executed"), h2.index("Thread Instructions Executed")
        ts, ti = sum(fnum(r[iS]) for _, r in rows) or 1.0, sum(fnum(r[iI]) for _, r in rows) or 1.0
        tt = sum(fnum(r[iT]) for _, r in rows)
        print(f"\n## Source attribution, launch {pick}: {ti:.4g} warp instructions, {tt:.4g} thread instructions, "
              f"{ts:.0f} stall samples\n")
        byfile = defaultdict(lambda: [0.0, 0.0, 0.0])
        for fn, r in rows:
            b = byfile[fn]
            b[0] += fnum(r[iS]); b[1] += fnum(r[iI]); b[2] += fnum(r[iT])
        print("| file | samples | warp instructions | threads / instruction |\n|---|---|---|---|")
        for fn, b in sorted(byfile.items(), key=lambda x: -x[1][0]):
            print(f"| {fn} | {b[0] / ts:.1%} | {b[1] / ti:.1%} | {b[2] / max(b[1], 1):.1f} |")
        print(f"\n### Hottest {args.top} source lines (by stall samples)\n")
        print("| samples | instr | thr/instr | location | source |\n|---|---|---|---|---|")
        for fn, r in sorted(rows, key=lambda x: -fnum(x[1][iS]))[:args.top]:
            text = r[1].strip().replace("|", "\\|")[:110]
            print(f"| {fnum(r[iS]) / ts:.1%} | {fnum(r[iI]) / ti:.1%} | {fnum(r[iT]) / max(fnum(r[iI]), 1):.1f} | "
                  f"{fn}:{r[0]} | `{text}` |")


if __name__ == "__main__":
    sys.exit(main())
