#!/bin/bash
# session 5: column-walk ROIAlign -- tests + microbench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -k roialign > gpurun_out/pytest_roi.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_roi.log | cut -c1-250
timeout 600 python tools/micro_post.py --old --out gpurun_out/micro_post_b32.json 2>&1 | tail -14
