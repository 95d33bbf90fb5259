#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_tc.py -m gpu -q --tb=short -x > gpurun_out/pytest_conv.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_conv.log | cut -c1-300
timeout 300 python tools/conv_bench.py --batch 16 --only osa5 2>&1 | tail -3
CM2_TC_TRIM=0 timeout 300 python tools/conv_bench.py --batch 16 --only osa5 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline --layers gpurun_out/layers_b16.txt > gpurun_out/bench_b16.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_b16.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['clocks'])"
grep -i "osa5\|fpn_inner5\|^p5\|p6\|p7" gpurun_out/layers_b16.txt
