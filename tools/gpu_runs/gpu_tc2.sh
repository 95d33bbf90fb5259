#!/bin/bash
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m pytest tests/test_gpu_conv_tc.py -q --tb=line -x > gpurun_out/pytest_tc_$name.log 2>&1
  echo "$name exit $?"; tail -4 gpurun_out/pytest_tc_$name.log | cut -c1-300
}
run auto CM2_TC_VARIANT=0
run v2plain CM2_TC_VARIANT=2
run v2merge_d0 CM2_TC_VARIANT=3 CM2_TC_DESC_MODE=0
run v2merge_d1 CM2_TC_VARIANT=3 CM2_TC_DESC_MODE=1
for v in 1 2 3; do
  CM2_TC_VARIANT=$v timeout 600 python bench.py --precision bf16 --batch 8 --steps 5 --warmup 3 --no-cpu-baseline --layers gpurun_out/layers_b8_variant$v.txt > gpurun_out/bench_variant$v.log 2>&1
  echo "bench variant $v exit $?"; tail -1 gpurun_out/bench_variant$v.log | cut -c1-200
done
