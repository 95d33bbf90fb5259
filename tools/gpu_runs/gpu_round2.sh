#!/bin/bash
# bench + profile artefacts (1 GPU).  gpurun_out is capped at 64 MiB: .ncu-rep files are reduced to CSV / text on the box.
mkdir -p gpurun_out
timeout 900 python bench.py --layers gpurun_out/layers_b16.txt > gpurun_out/bench_default.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_default.log | cut -c1-300
timeout 600 python tools/trace_step.py --batch 16 --steps 3 --graph 1 --out gpurun_out/trace_b16.txt > gpurun_out/trace.log 2>&1; echo "trace exit $?"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/ncu_step_b16.csv python bench.py --profile-step > gpurun_out/ncu_step.log 2>&1; echo "ncu step exit $?"
cap() {  # name, kernel regex, skip, count
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -f -o /tmp/$1 python bench.py --profile-step > gpurun_out/ncu_$1.log 2>&1; echo "ncu $1 exit $?"
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/ncu_$1_raw.csv 2>/dev/null
  if [ "$4" = "1" ]; then python tools/ncu_hot.py /tmp/$1.ncu-rep 40 > gpurun_out/ncu_$1_hot.txt 2>&1; fi
}
cap fcos_convs "conv_tc" 47 10
cap stem_osa2 "conv_tc" 0 9
cap fcos_tower1 "conv_tc" 49 1
cap gn_apply "gn_seg_apply" 1 1
cap ese_pool "ese_apply_pool" 0 1
cap paste "paste_masks" 0 1
cap im2col "im2col" 0 1
cap decode "fcos_decode" 0 1
cap roialign "roialign" 0 1
cap nms "nms_image" 0 1
cap mask_predict "mask_predict" 0 1
cp /tmp/fcos_tower1.ncu-rep gpurun_out/ 2>/dev/null
du -sh gpurun_out
