#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -k roialign > gpurun_out/pytest_roi.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_roi.log | cut -c1-250
timeout 600 python tools/micro_roi.py --knobs 0 2>&1 | tail -4
timeout 600 python tools/micro_roi.py --variants 2 --sort 2>&1 | tail -1
k=roialign_col
CM2_MICRO_EAGER=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/ncu_$k python tools/micro_post.py > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k exit $?"
