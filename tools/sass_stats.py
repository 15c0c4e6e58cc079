#!/usr/bin/env python
"""Static SASS statistics of the built kernels (runs here, no GPU): instructions per kernel, per source file
and line (nvdisasm -g line info), opcode histogram, spill instructions.

    python tools/sass_stats.py [--kernel REGEX] [--lines N] [--md profiles/rN_sass.md]

Used to see where a kernel's code comes from before spending GPU time, and to commit an opcode histogram
next to the ncu summaries (profiles/).
"""
import argparse
import collections
import re
import subprocess
import sys
import tempfile
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
OBJ = REPO / "python-raytracer_b200" / "csrc" / "sp_kernels.o"


def disassemble(obj):
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", str(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
        cubins = list(Path(tmp).glob("*.cubin"))
        return subprocess.run(["nvdisasm", "-g", "-c", str(cubins[0])], check=True, capture_output=True, text=True).stdout


def parse(text):
    """-> {kernel: [(file, line, opcode)]}"""
    out, cur_k, cur_loc = collections.OrderedDict(), None, ("?", 0)
    for line in text.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
        if m:
            cur_k = m.group(1); out[cur_k] = []; cur_loc = ("?", 0); continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur_loc = (m.group(1).split("/")[-1], int(m.group(2))); continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur_k:
            out[cur_k].append((cur_loc[0], cur_loc[1], m.group(1)))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kernel", default="sp_path_kernel")
    ap.add_argument("--lines", type=int, default=30)
    ap.add_argument("--md")
    ap.add_argument("--obj", default=str(OBJ))
    args = ap.parse_args()
    kernels = parse(disassemble(args.obj))
    lines = []
    for k, ins in kernels.items():
        if not re.search(args.kernel, k):
            continue
        ops = collections.Counter(o.split(".")[0] for _, _, o in ins)
        files = collections.Counter(f for f, _, _ in ins)
        locs = collections.Counter((f, l) for f, l, _ in ins)
        spill = sum(1 for _, _, o in ins if o.startswith(("STL", "LDL")))
        generic = sum(1 for _, _, o in ins if re.match(r"(LD|ST)(\.|$)", o))
        lines.append(f"## `{k}`\n")
        lines.append(f"{len(ins)} instructions ({len(ins) * 16 / 1024:.1f} KB), {spill} local-memory (spill / stack) instructions, "
                     f"{generic} generic LD/ST\n")
        lines.append("| opcode | count | | file | instructions |")
        lines.append("|---|---|---|---|---|")
        fo = files.most_common(12)
        for i, (op, c) in enumerate(ops.most_common(24)):
            f = f"| {fo[i][0]} | {fo[i][1]} |" if i < len(fo) else "| | |"
            lines.append(f"| {op} | {c} | {f}")
        lines.append("")
        lines.append("| instructions | source line |")
        lines.append("|---|---|")
        for (f, l), c in locs.most_common(args.lines):
            lines.append(f"| {c} | {f}:{l} |")
        lines.append("")
    text = "\n".join(lines)
    if args.md:
        Path(args.md).write_text("# Static SASS statistics (tools/sass_stats.py)\n\n" + text + "\n")
    print(text)


if __name__ == "__main__":
    sys.exit(main())
