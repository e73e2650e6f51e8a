"""Shared-memory wavefronts of the LSU by opcode from an `ncu --page source --csv` dump (the async-proxy operand reads of
UMMA / TMA are not in these counters; ncu_report_summary.py prints the kernel-wide L1 throughput beside them):
   ncu -i prof.ncu-rep --page source --csv > src.csv ; python tools/ncu_smem_wavefronts.py src.csv"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]


def num(r, h):
    try:
        return float(r[col[h]])
    except (ValueError, KeyError):
        return 0.0


wf, ideal, inst = Counter(), Counter(), Counter()
for r in data:
    w = num(r, "L1 Wavefronts Shared")
    if not w:
        continue
    src = r[col["Source"]].split()
    op = src[1] if src[0].startswith("@") else src[0]
    wf[op] += w
    ideal[op] += num(r, "L1 Wavefronts Shared Ideal")
    inst[op] += num(r, "Instructions Executed")
print(f"kernel: {rows[0][1] if rows[0][0] == 'Kernel Name' else '?'}")
print(f"{'opcode':12s} {'wavefronts':>12s} {'ideal':>12s} {'warp instr':>12s}")
for op, w in wf.most_common():
    print(f"{op:12s} {int(w):12d} {int(ideal[op]):12d} {int(inst[op]):12d}")
print(f"{'total':12s} {int(sum(wf.values())):12d} {int(sum(ideal.values())):12d}")
