#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -x -k roialign > gpurun_out/pytest_roi.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/pytest_roi.log | cut -c1-250
timeout 120 python tools/micro_roi.py --variants 3,2 2>&1 | tail -4
