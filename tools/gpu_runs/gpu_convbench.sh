#!/bin/bash
mkdir -p gpurun_out
for v in 1 3; do
  CM2_TC_VARIANT=$v timeout 300 python tools/conv_bench.py --batch 8 > gpurun_out/convbench_v$v.txt 2>&1; echo "convbench v$v exit $?"
  cat gpurun_out/convbench_v$v.txt
done
for v in 1 3; do
  CM2_TC_VARIANT=$v timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -o gpurun_out/ncu_fcos_v$v -f python tools/conv_bench.py --batch 8 --only fcos_tower_p3 --reps 1 > gpurun_out/ncu_fcos_v$v.log 2>&1; echo "ncu fcos v$v exit $?"
  CM2_TC_VARIANT=$v timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -o gpurun_out/ncu_osa2_v$v -f python tools/conv_bench.py --batch 8 --only osa2_3x3 --reps 1 > gpurun_out/ncu_osa2_v$v.log 2>&1; echo "ncu osa2 v$v exit $?"
done
