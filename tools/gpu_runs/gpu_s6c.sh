#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/micro_keypoints.txt
timeout 500 python -m pytest tests/test_gpu_keypoints.py tests/test_gpu_model.py -m gpu -q --tb=short -k "keypoint" 2>&1 | tail -15 | cut -c1-300 | tee gpurun_out/pytest_keypoints.log
for V in 1 0; do CM2_KP_VARIANT=$V timeout 120 python tools/micro_kp.py | tail -1 | tee -a gpurun_out/micro_keypoints.txt; done
