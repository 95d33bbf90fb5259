#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -q --tb=short -x -k bf16 -s 2>&1 | grep -E "bf16|passed|failed" | cut -c1-200
for L in stem1 stem2_phase stem3 osa2_3x3 osa5_3x3 mask_deconv fcos_regctr; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 1 -f -o gpurun_out/ncu_$L python tools/conv_bench.py --batch 16 --only $L > gpurun_out/ncu_$L.log 2>&1; echo "ncu $L exit $?"
done
ls -la gpurun_out/*.ncu-rep
