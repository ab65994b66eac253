"""Summarise an .ncu-rep: key raw metrics + stall samples by opcode and hottest instructions.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [n_hot]"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
nhot = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
        "sm__warps_active.avg.pct", "launch__registers_per_thread", "launch__grid_size", "sm__throughput.avg.pct",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum.pct", "sm__inst_executed_pipe_fp64.sum.pct",
        "sm__inst_executed_pipe_lsu.sum.pct", "sm__inst_executed_pipe_tensor", "sm__pipe_tensor", "smsp__issue_active.avg.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.max", "lts__t_bytes.sum", "sm__inst_executed_pipe_alu.sum.pct", "sm__inst_executed_pipe_xu.sum.pct",
        "sm__pipe_fp64_cycles_active", "sm__pipe_fma_cycles_active", "sm__pipe_alu_cycles_active")
for h, u, v in zip(hdr, units, vals):
    if h.startswith(want):
        print(f"{h:75s} {u:12s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot = sum(int(r[isamp]) for r in data)
c, ce = Counter(), Counter()
for r in data:
    t = r[ia].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    c[op] += int(r[isamp])
    ce[op] += int(r[iex])
print(f"\ntotal stall samples {tot}; warp instructions {sum(ce.values())}")
for op, n in c.most_common(18):
    print(f"  {op:10s} {100 * n / tot:5.1f}% samples  {ce[op]:>12d} instr")
print("\nhottest instructions:")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
for idx in sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:nhot]:
    r = data[idx]
    st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    print(f"  {int(r[isamp]):7d} {100 * int(r[isamp]) / tot:5.1f}%  ex {int(r[iex]):>9d}  #{idx:5d} {r[ia][:70]:70s} {st}")
