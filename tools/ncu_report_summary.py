"""Text summary of one or more .ncu-rep captures (the counters the DESIGN / VERDICT discussion uses):
   python tools/ncu_report_summary.py title:path.ncu-rep [title:path ...] > profiles/rNN_xxx.txt"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.max", "SM cycles"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers / thread"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/shared-memory throughput % (active)"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "LSU shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "LSU shared-memory bank conflicts"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe cycles active %"),
    ("sm__inst_executed.sum.per_cycle_active", "warp instructions / cycle (all SMs)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        res.append({h: (vals[i], units[i]) for i, h in enumerate(hdr)})
    return res


def stalls(path, top=6):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return ""
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = {h: 0.0 for h in cols}
    n = 0.0
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        for h in cols:
            try:
                tot[h] += float(r[col[h]])
            except ValueError:
                pass
        try:
            n += float(r[col["# Samples"]])
        except ValueError:
            pass
    best = sorted(tot.items(), key=lambda kv: -kv[1])[:top]
    return ", ".join(f"{k[6:]} {v / max(n, 1) * 100:.1f}%" for k, v in best)


for arg in sys.argv[1:]:
    title, path = arg.split(":", 1)
    for k in raw(path):
        name = k.get("Kernel Name", ("?", ""))[0]
        print(f"== {title}\n   kernel: {name}")
        for key, label in WANT:
            if key in k:
                v, u = k[key]
                print(f"   {label:44s} {v} {u}")
        s = stalls(path)
        if s:
            print(f"   warp stall samples (all warps)               {s}")
        print()
