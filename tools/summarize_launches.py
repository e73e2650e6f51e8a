"""Turn the ncu launch list of `bench.py` into the per-round profile files bench.py reads.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -c 600 --csv --log-file gpurun_out/bench_launches.csv python bench.py --steps 2 --warmup 1 --skip-diffusion --skip-cpu
    python tools/summarize_launches.py gpurun_out/bench_launches.csv profiles/r01_bench_launches
    python tools/summarize_launches.py gpurun_out/sampler_step_ncu.csv profiles/r02_sampler_step_launches sampler
        (launch list of `tools/forward_once.py sampler 32 16 fp16`: the last complete sampler timestep)

Writes <out>.csv (one bench step, one line per launch, with a share-by-kernel header) and <out>_summary.json
(`tensor_core_kernels.dram_bytes_per_launch` is what bench.py reports as roofline.traffic).  Per-launch times under
ncu are cold-cache, serialised and at burst clocks: compare SHARES of the step, not absolute times.
"""
import csv
import json
import re
import sys
from collections import OrderedDict

TENSOR = ("conv3x3_slab_kernel", "conv3x3_slab2_kernel", "conv3x3_chain_kernel", "igemm_kernel", "dense_block_kernel")


def short(name: str) -> str:
    name = re.sub(r"b200dn::|igemm::|<unnamed>::|\(anonymous namespace\)::|^void ", "", name)
    name = re.sub(r"\(KParams\)", "", name)
    name = name.replace("(int)", "").replace("(bool)", "")
    return name.split("(const")[0].strip()[:110]


def main() -> None:
    src, out = sys.argv[1], sys.argv[2]
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("==")) if len(r) >= 15]
    hdr, rows = rows[0], rows[1:]
    col = {h: i for i, h in enumerate(hdr)}
    launches = OrderedDict()
    for r in rows:
        lid = int(r[col["ID"]])
        d = launches.setdefault(lid, {"id": lid, "kernel": short(r[col["Kernel Name"]]),
                                      "grid": r[col["Grid Size"]].replace(", ", "x").strip("()")})
        v = float(r[col["Metric Value"]].replace(",", ""))
        m = r[col["Metric Name"]]
        if m == "gpu__time_duration.sum":
            d["us"] = v / 1e3 if r[col["Metric Unit"]] in ("ns", "nsecond") else v
        elif m == "dram__bytes_read.sum":
            d["rd"] = v
        elif m == "dram__bytes_write.sum":
            d["wr"] = v
    seq = list(launches.values())
    sampler = len(sys.argv) > 3 and sys.argv[3] == "sampler"
    if sampler:
        # one sampler timestep = the launches after a sampler_step_kernel up to and including the next one
        ends = [i for i, d in enumerate(seq) if "sampler_step" in d["kernel"]]
        if len(ends) < 2:
            raise SystemExit("need at least two sampler timesteps in the capture")
        step = seq[ends[-2] + 1:ends[-1] + 1]
    else:
        # one bench step = from a gauss_noise launch up to (not including) the next one.  Steps of the e2e leg may carry
        # a second forward (CUDA-graph warm-up), so take the first complete step with the smallest launch count.
        starts = [i for i, d in enumerate(seq) if "gauss_noise" in d["kernel"]]
        if len(starts) < 2:
            raise SystemExit("need at least two bench steps in the capture")
        segs = [seq[a:b] for a, b in zip(starts[:-1], starts[1:])]
        fewest = min(len(g) for g in segs)
        step = next(g for g in segs if len(g) == fewest)
    total_us = sum(d.get("us", 0.0) for d in step)
    by = OrderedDict()
    for d in step:
        b = by.setdefault(d["kernel"], {"us": 0.0, "launches": 0, "rd": 0.0, "wr": 0.0})
        b["us"] += d.get("us", 0.0)
        b["launches"] += 1
        b["rd"] += d.get("rd", 0.0)
        b["wr"] += d.get("wr", 0.0)
    tens = [d for d in step if any(t in d["kernel"] for t in TENSOR)]
    t_us = sum(d["us"] for d in tens)
    t_bytes = sum(d.get("rd", 0.0) + d.get("wr", 0.0) for d in tens)
    cmd = "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none python bench.py --steps 2 --warmup 1 --skip-diffusion --skip-cpu"
    what = "one bench step: RDUNet(128) bf16, B=64 x 256x256, noise + forward + PSNR/SSIM."
    if sampler:
        cmd = "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none python tools/forward_once.py sampler 32 16 fp16"
        what = ("ONE sampler timestep of DiffusionModel(RDUNet_T(32), 20).improved_sampling at B = 16: one forward over 32 images "
                "(ingest conv + the plan's tensor-core launches) + sampler_step.")
    with open(out + ".csv", "w") as f:
        f.write(f"# {cmd}\n# {what}  Cold-cache, serialised, "
                "burst clocks: compare SHARES.\n# share by kernel (us, %, launches, DRAM read MB, DRAM write MB):\n")
        for k, b in sorted(by.items(), key=lambda kv: -kv[1]["us"]):
            f.write(f"#   {k:58s} {b['us']:10.1f} us {100 * b['us'] / total_us:5.1f}%  n={b['launches']:3d}  rd {b['rd'] / 1e6:10.1f}  wr {b['wr'] / 1e6:10.1f}\n")
        f.write(f"#   total {total_us:.1f} us\nid,kernel,grid,us,dram_read_MB,dram_write_MB\n")
        for d in step:
            f.write(f"{d['id']},{d['kernel']},({d['grid']}),{d.get('us', 0):.1f},{d.get('rd', 0) / 1e6:.1f},{d.get('wr', 0) / 1e6:.1f}\n")
    summary = {
        "command": cmd + " (under ncu)", "step_total_us": total_us,
        "tensor_core_kernels": {"launches": len(tens), "us": t_us, "share_of_step": t_us / total_us,
                                "dram_bytes_per_step": t_bytes, "dram_bytes_per_launch": t_bytes / max(1, len(tens))},
        "by_kernel": {k: {"us": b["us"], "share": b["us"] / total_us, "launches": b["launches"],
                          "dram_read_MB": b["rd"] / 1e6, "dram_write_MB": b["wr"] / 1e6} for k, b in by.items()},
    }
    with open(out + "_summary.json", "w") as f:
        json.dump(summary, f, indent=1)
    print(json.dumps(summary["tensor_core_kernels"]))


if __name__ == "__main__":
    main()
