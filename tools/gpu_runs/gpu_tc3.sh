#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 600 python -m pytest tests/test_gpu_conv_tc.py -q --tb=line -x > gpurun_out/pytest_tc_$name.log 2>&1
  echo "$name exit $?"; tail -3 gpurun_out/pytest_tc_$name.log | cut -c1-300; }
run auto CM2_TC_VARIANT=0
run v1 CM2_TC_VARIANT=1
run v2plain CM2_TC_VARIANT=2
run v2merge CM2_TC_VARIANT=3
for v in 1 3 0; do
  CM2_TC_VARIANT=$v timeout 300 python tools/conv_bench.py --batch 8 > gpurun_out/convbench_v$v.txt 2>&1; echo "convbench v$v exit $?"; cat gpurun_out/convbench_v$v.txt
done
CM2_TC_VARIANT=0 timeout 600 python bench.py --precision bf16 --batch 8 --steps 5 --warmup 3 --no-cpu-baseline --layers gpurun_out/layers_b8_auto.txt > gpurun_out/bench_auto.log 2>&1
echo "bench exit $?"; tail -1 gpurun_out/bench_auto.log | cut -c1-250
