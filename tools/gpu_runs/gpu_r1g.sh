#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 600 python tools/micro_post.py --out gpurun_out/micro_post_b32.json 2>&1 | tail -8
timeout 600 python bench.py --no-cpu-baseline --layers gpurun_out/layers_b16.txt > gpurun_out/bench_b16.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_b16.log | cut -c1-300
timeout 600 python tools/trace_step.py --batch 16 --steps 3 --graph 1 --out gpurun_out/trace_b16.txt > gpurun_out/trace.log 2>&1; echo "trace exit $?"; sed -n 2,22p gpurun_out/trace_b16.txt
