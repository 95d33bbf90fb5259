#!/bin/bash
mkdir -p gpurun_out
for S in 0 1; do echo "== spin $S"; CM2_TC_SPIN=$S timeout 300 python tools/conv_bench.py --batch 16 2>&1 | tail -18; done | tee gpurun_out/convbench_spin.txt
for L in stem2_3x3 fcos_tower_p3; do for D in 15; do echo -n "$L spin1 dbg $D: "; CM2_TC_SPIN=1 CM2_TC_DEBUG=$D timeout 120 python tools/conv_bench.py --batch 16 --only $L 2>&1 | tail -1; done; done
