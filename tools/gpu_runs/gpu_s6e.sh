#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/micro_keypoints.txt
timeout 300 python -m pytest tests/test_gpu_keypoints.py tests/test_gpu_model.py -m gpu -q --tb=short -k "keypoint" 2>&1 | tail -15 | cut -c1-300 | tee gpurun_out/pytest_keypoints.log
for V in 1 0; do CM2_KP_VARIANT=$V timeout 120 python tools/micro_kp.py | tail -1 | tee -a gpurun_out/micro_keypoints.txt; done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:keypoints_decode --launch-skip 3 -c 1 -f -o gpurun_out/ncu_keypoints_decode python tools/micro_kp.py --iters 2 > gpurun_out/ncu_keypoints_decode.log 2>&1; echo "ncu kp exit $?"
