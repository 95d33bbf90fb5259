#!/bin/bash
# accuracy / throughput of the split-precision engine against the main-accumulator chunk length (CM2_TC_CHUNK)
mkdir -p gpurun_out
for C in "$@"; do
  CM2_TC_CHUNK=$C python tools/parity_report.py gpurun_out/parity_chunk$C.json fp32 2>&1 | tail -1 | cut -c1-420
  CM2_TC_CHUNK=$C python bench.py --precision fp32 --no-cpu-baseline --steps 10 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('chunk $C: value %.1f img/s  e2e %.1f  conv %.1f TFLOP/s' % (d['value'], d['e2e']['value'], d['roofline']['achieved']))"
done
