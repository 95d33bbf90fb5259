#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_tc.py -q --tb=short -x > gpurun_out/pytest_tc.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_tc.log
tail -40 gpurun_out/pytest_tc.log
if grep -q "pytest exit 0" gpurun_out/pytest_tc.log; then
  timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -q --tb=short > gpurun_out/pytest_rest.log 2>&1; echo "exit $?" >> gpurun_out/pytest_rest.log; tail -5 gpurun_out/pytest_rest.log
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -5 gpurun_out/smoke.log
  timeout 600 python bench.py --precision bf16 --batch 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_bf16.log; tail -3 gpurun_out/bench_bf16.log
fi
