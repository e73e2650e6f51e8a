"""Summarise an `ncu --page source --csv` dump: instruction and stall-sample totals, the hottest SASS lines.
   ncu -i prof.ncu-rep --page source --csv > src.csv ; python tools/ncu_source_summary.py src.csv [top=25]"""
import csv
import sys
from collections import Counter

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
kern = rows[0][1] if rows and rows[0][0] == "Kernel Name" else "?"
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]


def num(r, h):
    try:
        return float(r[col[h]])
    except ValueError:
        return 0.0


tot_inst = sum(num(r, "Instructions Executed") for r in data)
tot_samp = sum(num(r, "# Samples") for r in data)
print(f"kernel: {kern}\nSASS lines {len(data)}, warp instructions executed {tot_inst:.0f}, stall samples {tot_samp:.0f}")
st = Counter()
for r in data:
    for h in stall_cols:
        st[h] += num(r, h)
print("stall reasons (all samples): " + ", ".join(f"{k[6:]} {v / max(tot_samp, 1) * 100:.1f}%" for k, v in st.most_common(9)))
ops = Counter()
for r in data:
    op = r[col["Source"]].split()[0] if r[col["Source"]].split() else "?"
    if op.startswith("@"):
        op = r[col["Source"]].split()[1]
    ops[op.split(".")[0]] += num(r, "Instructions Executed")
print("instructions by opcode: " + ", ".join(f"{k} {v / max(tot_inst, 1) * 100:.1f}%" for k, v in ops.most_common(14)))
print(f"-- top {top} lines by stall samples")
for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:top]:
    dom = max(stall_cols, key=lambda h: num(r, h))
    print(f"  {r[col['Address']][-5:]} samp {num(r, '# Samples'):7.0f} ({num(r, '# Samples') / max(tot_samp, 1) * 100:4.1f}%) inst {num(r, 'Instructions Executed'):9.0f} "
          f"{dom[6:]:14s} {r[col['Source']][:90]}")
