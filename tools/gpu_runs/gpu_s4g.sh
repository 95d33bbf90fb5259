#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --tb=short -x -k "stream or postprocessed" > gpurun_out/pytest_stream.log 2>&1; echo "pytest exit $?"
tail -15 gpurun_out/pytest_stream.log | cut -c1-250
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_b16.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_b16.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], json.dumps(d['e2e']), d['clocks'])"
