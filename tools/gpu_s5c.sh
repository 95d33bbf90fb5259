#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/micro_roi.py --variants 2,1 --knobs 0,7,100,107 2>&1 | tail -10
