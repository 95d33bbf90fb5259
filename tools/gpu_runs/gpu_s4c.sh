#!/bin/bash
# session-4: microbench with graph timing + ncu full captures of the new post-process kernels
mkdir -p gpurun_out
timeout 600 python tools/micro_post.py --old --out gpurun_out/micro_post_b32.json 2>&1 | tail -12
for k in paste_fused roialign_roi fcos_decode_tiles spatial_attention_smem; do
  CM2_MICRO_EAGER=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/ncu_$k python tools/micro_post.py > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k exit $?"
done
