#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_kernels.py tests/test_gpu_model.py -q --tb=short -x > gpurun_out/pytest_seg.log 2>&1; echo "pytest exit $?"
tail -25 gpurun_out/pytest_seg.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --batch 8 --steps 5 --warmup 3 --no-cpu-baseline --layers gpurun_out/layers_b8_seg.txt > gpurun_out/bench_seg_b8.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_seg_b8.log | cut -c1-200
timeout 600 python bench.py --batch 16 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_seg_b16.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_seg_b16.log | cut -c1-200
