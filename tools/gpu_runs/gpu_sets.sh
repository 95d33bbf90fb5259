#!/bin/bash
mkdir -p gpurun_out
for C in "2 1" "4 1" "2 2" "4 2"; do set -- $C; echo "== sets v1=$1 v2=$2"; CM2_TC_EPI_SETS_V1=$1 CM2_TC_EPI_SETS_V2=$2 timeout 300 python tools/conv_bench.py --batch 16 2>&1 | tail -18 | awk '{printf "%-22s %s %s\n", $1, $4, $6}'; done | tee gpurun_out/convbench_sets.txt
timeout 300 python -m pytest tests/test_gpu_conv_tc.py -q -x 2>&1 | tail -2
CM2_TC_EPI_SETS_V1=4 CM2_TC_EPI_SETS_V2=2 timeout 300 python -m pytest tests/test_gpu_conv_tc.py -q -x 2>&1 | tail -2
