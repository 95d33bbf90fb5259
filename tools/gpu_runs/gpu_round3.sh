#!/bin/bash
# round-end style validation + profile artefacts (1 GPU).  gpurun_out is capped at 64 MiB: .ncu-rep files are reduced on the box.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log | cut -c1-200
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --layers gpurun_out/layers_b16.txt > gpurun_out/bench_default.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_default.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "ref exit $?"
timeout 600 python tools/trace_step.py --batch 16 --steps 3 --graph 1 --out gpurun_out/trace_b16.txt > gpurun_out/trace.log 2>&1; echo "trace exit $?"
timeout 600 python tools/micro_post.py --out gpurun_out/micro_post_b32.json > gpurun_out/micro.log 2>&1; echo "micro exit $?"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/ncu_step_b16.csv python bench.py --profile-step > gpurun_out/ncu_step.log 2>&1; echo "ncu step exit $?"
cap() {  # name, kernel regex, skip, count
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -f -o /tmp/$1 python bench.py --profile-step > gpurun_out/ncu_$1.log 2>&1; echo "ncu $1 exit $?"
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/ncu_$1_raw.csv 2>/dev/null
  if [ "$4" = "1" ]; then python tools/ncu_hot.py /tmp/$1.ncu-rep 40 > gpurun_out/ncu_$1_hot.txt 2>&1; fi
}
cap all_convs "conv_tc" 0 69
cap fcos_tower1 "conv_tc" 49 1
cap gn_apply "gn_seg_apply" 1 1
cap ese_pool "ese_apply_pool" 0 1
cap paste "paste_window" 0 1
du -sh gpurun_out
