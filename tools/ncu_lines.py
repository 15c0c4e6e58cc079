#!/usr/bin/env python
"""Rank the source lines of one captured launch by warp instructions executed (ncu source page, read offline).

    python tools/ncu_lines.py report.ncu-rep [--kernel-regex sp_level] [--kernel-index 1] [--top 60] [--rays N]
"""
import argparse, csv, io, subprocess
from collections import defaultdict


def fnum(x):
    try:
        return float(str(x).replace(",", ""))
    except ValueError:
        return 0.0


ap = argparse.ArgumentParser()
ap.add_argument("report"); ap.add_argument("--kernel-regex", default="sp_level"); ap.add_argument("--kernel-index", type=int, default=1)
ap.add_argument("--top", type=int, default=60); ap.add_argument("--rays", type=float, default=0.0)
a = ap.parse_args()
src = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id",
                      f"::regex:{a.kernel_regex}:{a.kernel_index}"], check=True, capture_output=True, text=True).stdout
rows, cur, h2 = [], None, None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        h2 = r
    elif h2 and len(r) == len(h2) and r[2] == "-":
        rows.append((cur, r))
iS, iI, iT = h2.index("# Samples"), h2.index("Instructions Executed"), h2.index("Thread Instructions Executed")
ti = sum(fnum(r[iI]) for _, r in rows); ts = sum(fnum(r[iS]) for _, r in rows)
per = (lambda x: f"{x / a.rays * 32:7.1f}") if a.rays else (lambda x: f"{x / ti:6.1%}")
print(f"total warp instr {ti:.4g}, samples {ts:.0f}" + (f", warp instr per 32 rays {ti / a.rays * 32:.0f}" if a.rays else ""))
for fn, r in sorted(rows, key=lambda x: -fnum(x[1][iI]))[:a.top]:
    print(f"{per(fnum(r[iI]))} {fnum(r[iS]) / ts:6.1%} {fnum(r[iT]) / max(fnum(r[iI]), 1):5.1f}  {fn}:{r[0]:>4}  {r[1].strip()[:120]}")
