#!/bin/bash
# session 5: ncu full capture of a ROIAlign kernel (K = kernel name regex, V = CM2_ROIALIGN_VARIANT)
mkdir -p gpurun_out
k=${K:-roialign_col}
CM2_ROIALIGN_VARIANT=${V:-2} CM2_MICRO_EAGER=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/ncu_$k python tools/micro_post.py > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k exit $?"
