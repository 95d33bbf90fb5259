#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -x -k "paste or roialign or fcos_post" > gpurun_out/pytest_post.log 2>&1; echo "pytest exit $?"
tail -15 gpurun_out/pytest_post.log | cut -c1-250
timeout 600 python tools/micro_post.py --out gpurun_out/micro_post_b32.json 2>&1 | tail -12
for k in $NCU_KERNELS; do
  CM2_MICRO_EAGER=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/ncu_$k python tools/micro_post.py > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k exit $?"
done
