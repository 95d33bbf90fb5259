#!/usr/bin/env python
"""Per-kernel SASS opcode summary of libcm2.so: which entry points carry tcgen05 / TMEM / TMA instructions.

    python tools/sass_summary.py [--out profiles/r2_sass_opcodes.txt]

Counts, per kernel function of `cuobjdump -sass centermask2_b200/libcm2.so`: UTCHMMA / UTCQMMA (tcgen05.mma), UTMALDG
(TMA tensor load), UTMASTG (TMA store), LDTM / STTM (tcgen05.ld / .st), UTCBAR (tcgen05.commit -> mbarrier), UTCATOMSWS
(TMEM alloc), SYNCS (mbarrier ops), HMMA (mma.sync), plus the instruction total.  Mnemonics per
/opt/skills/guides/B200_PROFILING.md."""
import argparse
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "HMMA", "LDGSTS", "FFMA2"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=os.path.join(ROOT, "centermask2_b200", "libcm2.so"))
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    sass = subprocess.run(["cuobjdump", "-sass", args.lib], capture_output=True, text=True, check=True).stdout
    demangle = {}
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?PT?\d*\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            counts[cur]["total"] += 1
            if op in OPS:
                counts[cur][op] += 1
    names = list(counts)
    try:
        out = subprocess.run(["c++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
        demangle = dict(zip(names, out))
    except (OSError, subprocess.CalledProcessError):
        pass
    rows = ["# SASS opcode counts per kernel of libcm2.so (sm_100a); cuobjdump -sass; see tools/sass_summary.py",
            "{:78s} {:>7s} ".format("kernel", "instrs") + " ".join("{:>7s}".format(o[:7]) for o in OPS)]
    for n, c in sorted(counts.items(), key=lambda kv: -(kv[1]["UTCHMMA"] * 1000 + kv[1]["total"])):
        short = re.sub(r"\(.*", "", demangle.get(n, n)).replace("void ", "")
        rows.append("{:78s} {:7d} ".format(short[:78], c["total"]) + " ".join("{:7d}".format(c[o]) for o in OPS))
    txt = "\n".join(rows)
    print(txt)
    if args.out:
        with open(args.out, "w") as f:
            f.write(txt + "\n")


if __name__ == "__main__":
    main()
