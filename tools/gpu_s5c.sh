#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/micro_roi.py --knobs ${KNOBS:-0,1,2,3,9,10,11} 2>&1 | tail -12
