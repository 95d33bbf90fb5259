#!/bin/bash
mkdir -p gpurun_out
for L in stem1 stem2_3x3 stem3 osa2_3x3 osa2_cat fcos_regctr fcos_logits osa5; do
for D in 0 1 2 3 4 6 7; do echo -n "$L dbg $D: "; CM2_TC_DEBUG=$D timeout 120 python tools/conv_bench.py --batch 16 --only $L 2>&1 | tail -1; done; done | tee gpurun_out/convdbg_b16.txt
