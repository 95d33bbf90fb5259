#!/usr/bin/env python
"""Summarise an ncu CSV of ONE profiled step (bench.py --profile-step under
`ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`):

    python tools/ncu_step_summary.py gpurun_out/ncu_step_b16.csv --batch 16 --out profiles/r1_ncu_step_b16_summary.txt

Writes a per-kernel table (launches, time, share, DRAM bytes, achieved DRAM GB/s) and profiles/conv_traffic.json
(DRAM bytes of all convolution launches of the step), which bench.py reports as roofline.traffic.
Per-launch times under ncu are cold-cache and serialised: shares are meaningful, absolutes are not bench values.
"""
import argparse
import collections
import csv
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    lines = [l for l in open(args.csv) if not l.startswith("==")]
    per = collections.defaultdict(dict)
    names = {}
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        u = row["Metric Unit"]
        m = row["Metric Name"]
        if m.startswith("gpu__time"):
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)                     # -> us
        else:
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)         # -> bytes
        per[row["ID"]][m] = v
        names[row["ID"]] = re.sub(r"<.*|\(.*", "", row["Kernel Name"])
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for i, d in per.items():
        a = agg[names[i]]
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[2] += d.get("dram__bytes_read.sum", 0.0)
        a[3] += d.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    out = ["# one step, batch {} {}: per-kernel totals from ncu (cold-cache, serialised launches)".format(args.batch, args.precision),
           "# total kernel time {:.3f} ms over {} launches".format(tot / 1e3, sum(a[0] for a in agg.values())),
           "{:44s} {:>6s} {:>10s} {:>7s} {:>10s} {:>10s} {:>9s}".format("kernel", "n", "ms", "share", "rd MB", "wr MB", "DRAM GB/s")]
    conv_bytes = 0.0
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        gbs = (a[2] + a[3]) / (a[1] * 1e-6) / 1e9 if a[1] else 0.0
        out.append("{:44s} {:6d} {:10.3f} {:6.1f}% {:10.1f} {:10.1f} {:9.0f}".format(k[:44], a[0], a[1] / 1e3, 100 * a[1] / tot,
                                                                               a[2] / 1e6, a[3] / 1e6, gbs))
        if "conv_tc" in k or "conv_simt" in k or "stem1_fused" in k or "splitk_finish" in k:      # every launch bench.py times as a convolution
            conv_bytes += a[2] + a[3]
    txt = "\n".join(out)
    print(txt)
    if args.out:
        with open(args.out, "w") as f:
            f.write(txt + "\n")
    with open(os.path.join(ROOT, "profiles", "conv_traffic.json"), "w") as f:
        json.dump({"images_per_gpu": args.batch, "precision": args.precision, "conv_dram_bytes_per_step": conv_bytes,
                   "source": os.path.basename(args.csv) + " (ncu dram__bytes_read.sum + dram__bytes_write.sum over every conv launch of one step)"}, f)


if __name__ == "__main__":
    main()
