#!/bin/bash
mkdir -p gpurun_out
for v in 1 3; do for d in 0 1 2 4 3 6 7; do
  echo "== variant $v debug $d"
  CM2_TC_VARIANT=$v CM2_TC_DEBUG=$d timeout 300 python tools/conv_bench.py --batch 8 --only fcos_tower_p3 2>&1 | tail -1
  CM2_TC_VARIANT=$v CM2_TC_DEBUG=$d timeout 300 python tools/conv_bench.py --batch 8 --only osa2_3x3 2>&1 | tail -1
done; done | tee gpurun_out/convdbg.txt
