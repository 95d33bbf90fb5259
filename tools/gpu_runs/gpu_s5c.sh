#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/micro_roi.py --variants 3,2 --order 6:4:0,1:4:0,2:6:3,2:7:3,3:5:1,3:6:1,4:5:0,3:6:0 2>&1 | tail -16
