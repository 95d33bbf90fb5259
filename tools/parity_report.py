#!/usr/bin/env python
"""Write the full-size parity report (tests/fullsize.py) of every engine to a JSON file.

    python tools/parity_report.py gpurun_out/r2_parity_fullsize.json [precision ...]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from tests import fullsize      # noqa: E402


def main():
    out = sys.argv[1]
    precisions = sys.argv[2:] or ["fp32", "fp32_simt", "bf16"]
    reports = {}
    for p in precisions:
        try:
            reports[p] = fullsize.deviation_report(p)
        except Exception as e:          # keep the other engines' numbers
            reports[p] = {"error": repr(e)}
        print(p, json.dumps(reports[p]))
    with open(out, "w") as f:
        json.dump(reports, f, indent=1)


if __name__ == "__main__":
    main()
