#!/bin/bash
# session-4: new post-process kernels (fused paste, merged-tap ROIAlign, flat-tile decode): tests + microbench old vs new
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -k "paste or roialign or fcos_post" > gpurun_out/pytest_post.log 2>&1; echo "pytest exit $?"
tail -30 gpurun_out/pytest_post.log | cut -c1-250
timeout 600 python tools/micro_post.py --old --out gpurun_out/micro_post_b32.json 2>&1 | tail -12
