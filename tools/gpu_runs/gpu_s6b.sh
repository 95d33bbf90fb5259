#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/micro_kp.py --iters 5 | tail -1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:keypoints_decode --launch-skip 3 -c 1 -f -o gpurun_out/ncu_keypoints_decode python tools/micro_kp.py --iters 2 > gpurun_out/ncu_keypoints_decode.log 2>&1; echo "ncu exit $?"
