#!/bin/bash
mkdir -p gpurun_out
CM2_TC_DEBUG=15 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -f -o /tmp/skel_stem2 python tools/conv_bench.py --batch 16 --only stem2_3x3 > gpurun_out/ncu_skel.log 2>&1; echo "exit $?"
ncu -i /tmp/skel_stem2.ncu-rep --page source --csv --print-source sass > gpurun_out/skel_stem2_sass.csv 2>/dev/null
python tools/ncu_hot.py /tmp/skel_stem2.ncu-rep 30
