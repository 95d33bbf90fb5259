#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/conv_bench.py --batch 16 > gpurun_out/cb_default.txt 2>&1; echo "default exit $?"
CM2_TC_PAIR=1 CM2_TC_VARIANT=3 timeout 300 python tools/conv_bench.py --batch 16 > gpurun_out/cb_pair_v3.txt 2>&1; echo "pair v3 exit $?"
CM2_TC_PAIR=1 CM2_TC_VARIANT=3 CM2_TC_B_RESIDENT=0 timeout 300 python tools/conv_bench.py --batch 16 > gpurun_out/cb_pair_v3_nores.txt 2>&1; echo "pair v3 nores exit $?"
CM2_TC_VARIANT=3 timeout 300 python tools/conv_bench.py --batch 16 > gpurun_out/cb_v3.txt 2>&1; echo "v3 exit $?"
paste <(awk '{print $1, $6}' gpurun_out/cb_default.txt) <(awk '{print $6}' gpurun_out/cb_v3.txt) <(awk '{print $6}' gpurun_out/cb_pair_v3.txt) <(awk '{print $6}' gpurun_out/cb_pair_v3_nores.txt) | column -t
