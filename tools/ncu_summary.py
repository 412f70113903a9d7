#!/usr/bin/env python3
"""Summarise an .ncu-rep capture of the MPC kernel into a small tracked text file:
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/<name>.txt
Writes the key raw metrics, the executed-instruction mix and the warp-stall mix (from the source page)."""
import collections
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_per_inst_issued.ratio",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "lts__t_bytes.sum",
]
lines = [f"# ncu --set full --clock-control none --import-source on : {rep}", f"# kernel: {vals[hdr.index('Kernel Name')]}"]
for k in KEYS:
    if k in hdr:
        i = hdr.index(k)
        lines.append(f"{k:75s} {vals[i]:>18s} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ix = {c: i for i, c in enumerate(h)}
byop, stall, tot = collections.Counter(), collections.Counter(), 0
stall_op = collections.Counter()   # stall samples attributed to the opcode the warp was waiting to issue
scols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
for r in rows[2:]:
    if len(r) != len(h):
        continue
    toks = r[ix["Source"]].split()
    op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
    ex = int(r[ix["Instructions Executed"]])
    byop[op] += ex
    tot += ex
    for c in scols:
        stall[c] += int(r[ix[c]])
        stall_op[op] += int(r[ix[c]])
lines.append(f"\n# executed warp-instructions by opcode (total {tot})")
for op, c in byop.most_common(16):
    lines.append(f"{op:10s} {100.0 * c / tot:6.2f} %")
ts = sum(stall.values())
lines.append("\n# warp stall samples (all samples)")
for k, v in stall.most_common(9):
    lines.append(f"{k:24s} {100.0 * v / ts:6.2f} %")
lines.append("\n# warp stall samples by the opcode waiting to issue (where the latency is exposed)")
for op, c in stall_op.most_common(10):
    lines.append(f"{op:10s} {100.0 * c / max(ts, 1):6.2f} %")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
