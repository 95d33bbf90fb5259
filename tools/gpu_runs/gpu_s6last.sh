#!/bin/bash
# last check of the final tree: full GPU suite + one bench line (no CPU baseline leg)
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 100 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_b16_nocpu.log; python -c "
import json
d=json.loads(open('gpurun_out/bench_b16_nocpu.log').read())
print('value', round(d['value'],1), 'e2e stream', round(d['e2e']['value'],1), 'per-call', round(d['e2e']['forward_per_call']['value'],1))"
