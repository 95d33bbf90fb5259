#!/usr/bin/env python
"""Top SASS instructions of an ncu report by warp-stall samples: python tools/ncu_hot.py report.ncu-rep [N]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        samp = int(r[col["# Samples"]])
    except ValueError:
        continue
    data.append((samp, r))
total = sum(s for s, _ in data)
inst_total = sum(int(r[col["Instructions Executed"]] or 0) for _, r in data)
print("total samples", total, "instructions executed", inst_total)
for samp, r in sorted(data, key=lambda t: -t[0])[:top]:
    st = sorted(((int(r[col[s]] or 0), s) for s in stalls), reverse=True)[:2]
    print("{:6.2f}% {:>9s} inst  {:60s} {}".format(100.0 * samp / total, r[col["Instructions Executed"]], r[col["Source"]][:60],
                                                 " ".join("{}={}".format(s[6:], n) for n, s in st if n)))
