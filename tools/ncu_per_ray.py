#!/usr/bin/env python
"""Per-ray figures of a configuration from an ncu launch list + the bench line of the SAME command run without ncu.

    ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum \\
        --clock-control none --csv --log-file launches.csv python bench.py <args>
    python tools/ncu_per_ray.py launches.csv bench_line.json <config> [profiles/r2_per_ray.json]

Sums the metrics over the wavefront kernels of the run (level / warp / hit / shade / trace / shadow kernels) (all steps and warm-ups alike: every step traces the same rays)
and divides by the rays those launches traced: steps x rays_per_frame of the bench line (the ncu run executes the
same command, so the same number of frames).  bench.py reads the result for `roofline_issue` / `roofline_hbm.traffic`.
"""
import csv
import json
import sys
from collections import defaultdict
from pathlib import Path


def main():
    launches, bench, config = sys.argv[1], json.loads(Path(sys.argv[2]).read_text().strip().splitlines()[-1]), sys.argv[3]
    out_path = Path(sys.argv[4]) if len(sys.argv) > 4 else None
    rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
    hdr = rows[0]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    tot, n = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        name = r[ki].replace("void ", "").split("(")[0].split("<")[0]
        if not name.startswith(("sp_level", "sp_warp", "sp_hit", "sp_shade", "sp_trace", "sp_shadow")):
            continue
        tot[r[mi]] += float(r[vi].replace(",", ""))
        if r[mi] == "gpu__time_duration.sum":
            n[name] += 1
    frames = bench["steps"] + bench["warmup"]
    if bench.get("e2e"):
        frames += 1 + bench["e2e"]["steps"]
    rays = bench["config"]["rays_per_frame"] * frames
    entry = {
        "thread_inst_per_ray": tot["smsp__thread_inst_executed.sum"] / rays,
        "warp_inst_per_ray": tot["smsp__inst_executed.sum"] / rays,
        "dram_bytes_per_ray": (tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"]) / rays,
        "active_lanes_per_warp_inst": tot["smsp__thread_inst_executed.sum"] / max(tot["smsp__inst_executed.sum"], 1.0),
        "launches": dict(n), "frames": frames, "rays": rays,
        "command": "bench.py " + " ".join(sys.argv[5:]) if len(sys.argv) > 5 else None,
    }
    print(json.dumps(entry, indent=1))
    if out_path:
        data = json.loads(out_path.read_text()) if out_path.exists() else {}
        data[config] = entry
        out_path.write_text(json.dumps(data, indent=1) + "\n")


if __name__ == "__main__":
    main()
