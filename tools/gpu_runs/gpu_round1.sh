#!/bin/bash
# first GPU round: kernel tests, model parity, smoke, a short fp32 bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 -x --tb=short > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -15 gpurun_out/smoke.log
timeout 600 python bench.py --precision fp32 --batch 4 --steps 3 --warmup 3 > gpurun_out/bench_fp32.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_fp32.log
tail -5 gpurun_out/bench_fp32.log
