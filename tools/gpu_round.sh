#!/bin/bash
# The round's single-GPU evidence set: every bench line, the reference arm, the ncu launch list of one step.
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/r2_bench_$tag.log 2>&1; echo "[bench $tag] exit $?"; tail -1 gpurun_out/r2_bench_$tag.log > gpurun_out/r2_bench_$tag.json; cut -c1-260 gpurun_out/r2_bench_$tag.json; }
run v39_bf16 --layers gpurun_out/r2_layers_v39_bf16.txt
run v39_fp32 --precision fp32 --no-cpu-baseline --layers gpurun_out/r2_layers_v39_fp32.txt
run lite --config lite --no-cpu-baseline --layers gpurun_out/r2_layers_lite.txt
run v99 --config v99 --no-cpu-baseline --layers gpurun_out/r2_layers_v99.txt
run post --config post
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2>gpurun_out/r2_bench_reference.err; echo "[reference] exit $?"; cut -c1-300 gpurun_out/r2_bench_reference.json
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/r2_ncu_step_b16.csv python bench.py --profile-step > gpurun_out/r2_ncu_step.log 2>&1; echo "[ncu-step] exit $?"
python tools/ncu_step_summary.py gpurun_out/r2_ncu_step_b16.csv --out gpurun_out/r2_ncu_step_b16_summary.txt 2>&1 | tail -3
# programmatic dependent launch on / off, same box
for k in 0 1; do CM2_PDL=$k timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('CM2_PDL=$k', d['ms_per_step'], 'ms/step', d['value'], 'img/s')"; done | tee gpurun_out/r2_pdl_ab.txt
