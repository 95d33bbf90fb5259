#!/bin/bash
# device-resident ms/step of the headline config at small per-GPU batches and of the Lite config, with / without branch streams
mkdir -p gpurun_out
for BR in 1 0; do
  for ARGS in "--batch 16" "--batch 4" "--batch 2" "--config lite" "--config v99"; do
    CM2_BRANCH_STREAMS=$BR python bench.py $ARGS --no-cpu-baseline --no-soak --steps 30 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('branch_streams=$BR  %-14s value %7.0f img/s  %.3f ms/step  e2e %7.0f  conv %.0f TF/s (frac %.2f)  whole-step %.0f TF/s' % ('$ARGS', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['whole_step_tflops']))"
  done
done
