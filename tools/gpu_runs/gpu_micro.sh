#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/micro_post.py --out gpurun_out/micro_post_b32.json 2>&1 | tail -12
timeout 600 python tools/trace_step.py --batch 16 --steps 3 --graph 1 --out gpurun_out/trace_b16.txt > gpurun_out/trace.log 2>&1; echo "trace exit $?"; sed -n 2,12p gpurun_out/trace_b16.txt
