#!/usr/bin/env python3
"""Stall samples of an .ncu-rep capture attributed to CUDA source lines (needs -lineinfo and --import-source on):
    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top_n] [file-filter]
Prints, for the lines with the most stall samples: file:line, share of all samples, the dominant stall reasons,
executed warp-instructions, and the source text."""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
flt = sys.argv[3] if len(sys.argv) > 3 else ""
by_inst = len(sys.argv) > 4 and sys.argv[4] == "inst"   # sort by executed warp-instructions instead of stall samples
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
out, fname, h, ix, scols = [], "", None, None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        h = r
        ix = {}
        for i, c in enumerate(h):
            ix.setdefault(c, i)
        scols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
        continue
    if h is None or len(r) != len(h) or r[0] == "" or r[2] != "-":
        continue   # keep the per-source-line rows (they carry the totals of their SASS instructions)
    num = lambda v: int(v) if v not in ("-", "") else 0
    st = {c: num(r[ix[c]]) for c in scols}
    out.append((sum(st.values()), num(r[ix["Instructions Executed"]]), fname, r[0], r[1], st))
total = sum(r[0] for r in out) or 1
print(f"# {rep}: {total} stall samples over {len(out)} source lines")
byfile = collections.Counter()
for t, ex, f, ln, text, st in out:
    byfile[f] += t
print("# by file: " + ", ".join(f"{f} {100.0 * v / total:.1f}%" for f, v in byfile.most_common(6)))
tot_inst = sum(r[1] for r in out) or 1
for t, ex, f, ln, text, st in sorted([o for o in out if flt in o[2]], key=lambda r: -(r[1] if by_inst else r[0]))[:top]:
    reasons = ", ".join(f"{k.replace('stall_', '')} {100.0 * v / max(t, 1):.0f}%" for k, v in collections.Counter(st).most_common(3) if v)
    print(f"{100.0 * t / total:6.2f} %  inst {100.0 * ex / tot_inst:5.2f} %  {f}:{ln:>4s}  [{reasons}]  {text.strip()[:100]}")
