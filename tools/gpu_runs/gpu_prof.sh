#!/bin/bash
# per-layer table + ncu launch list + one full ncu capture of the tcgen05 conv kernel
mkdir -p gpurun_out
B=${1:-8}
timeout 600 python bench.py --precision bf16 --batch $B --steps 3 --warmup 3 --no-cpu-baseline --layers gpurun_out/layers_b$B.txt > gpurun_out/bench_prof.log 2>&1 || { tail -20 gpurun_out/bench_prof.log; exit 1; }
tail -1 gpurun_out/bench_prof.log | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_b$B.csv python bench.py --precision bf16 --batch $B --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 300 -c 3 -o gpurun_out/prof_conv_tc -f python bench.py --precision bf16 --batch $B --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out
