#!/bin/bash
# cta_group::2 pair kernel: correctness of the conv tests forced through it, then the conv microbench with / without pairs
mkdir -p gpurun_out
CM2_TC_PAIR=1 CM2_TC_VARIANT=3 CM2_TC_B_RESIDENT=0 timeout 300 python -m pytest tests/test_gpu_conv_tc.py -m gpu -q --tb=short -x -k "conv_tc_halo or multi_wave" > gpurun_out/pytest_pair.log 2>&1; echo "pytest pair exit $?"
tail -25 gpurun_out/pytest_pair.log | cut -c1-300
CM2_TC_PAIR=1 timeout 300 python tools/conv_bench.py --batch 16 --only osa > gpurun_out/convbench_pair.txt 2>&1; echo "bench pair exit $?"; cat gpurun_out/convbench_pair.txt | tail -20
timeout 300 python tools/conv_bench.py --batch 16 --only osa > gpurun_out/convbench_nopair.txt 2>&1; echo "bench nopair exit $?"; cat gpurun_out/convbench_nopair.txt | tail -20
