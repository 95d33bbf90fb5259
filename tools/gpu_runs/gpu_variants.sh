#!/bin/bash
mkdir -p gpurun_out
for V in 0 1 2 3; do CM2_TC_VARIANT=$V timeout 300 python tools/conv_bench.py --batch 16 > gpurun_out/convbench_b16_v$V.txt 2>&1; echo "variant $V exit $?"; done
paste gpurun_out/convbench_b16_v0.txt gpurun_out/convbench_b16_v1.txt gpurun_out/convbench_b16_v2.txt gpurun_out/convbench_b16_v3.txt | awk '{printf "%-20s auto %8s  v1 %8s  v2plain %8s  v2merge %8s\n", $1, $6, $14, $22, $30}'
