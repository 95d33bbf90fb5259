#!/bin/bash
# round-end style validation + profile artefacts (1 GPU)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --layers gpurun_out/layers_b16.txt > gpurun_out/bench_default.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_default.log | cut -c1-400
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "ref exit $?"; tail -1 gpurun_out/bench_reference.log | cut -c1-300
timeout 600 python tools/trace_step.py --batch 16 --steps 3 --graph 1 --out gpurun_out/trace_b16.txt > gpurun_out/trace.log 2>&1; echo "trace exit $?"; head -24 gpurun_out/trace_b16.txt
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/ncu_step_b16.csv python bench.py --profile-step > gpurun_out/ncu_step.log 2>&1; echo "ncu step exit $?"
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:"conv_tc" -c 69 -f -o gpurun_out/ncu_full_convs python bench.py --profile-step > gpurun_out/ncu_full_convs.log 2>&1; echo "ncu full convs exit $?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"gn_seg_apply|ese_apply_pool|paste_masks|fcos_decode|roialign|nms_image|mask_predict|im2col|spatial_att" -c 40 -f -o gpurun_out/ncu_full_misc python bench.py --profile-step > gpurun_out/ncu_full_misc.log 2>&1; echo "ncu full misc exit $?"
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv
