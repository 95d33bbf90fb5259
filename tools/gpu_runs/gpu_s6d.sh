#!/bin/bash
# does the nvidia-smi clock sampler (driver queries every 20 ms) slow the host-in-the-loop e2e regions?
mkdir -p gpurun_out
for MS in 20 100 1000; do
timeout 300 python bench.py --no-cpu-baseline --clock-ms $MS 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('clock-ms $MS: value', round(d['value'],1), 'e2e stream', round(d['e2e']['value'],1), 'per-call', round(d['e2e']['forward_per_call']['value'],1), 'clock samples', d['clocks']['samples'])" | tee -a gpurun_out/clock_sampler_effect.txt
done
