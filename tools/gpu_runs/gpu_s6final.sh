#!/bin/bash
# session-6 round-end style validation + profile artefacts (1 GPU)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python bench.py --layers gpurun_out/layers_b16.txt > gpurun_out/bench_b16.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_b16.log | cut -c1-300
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/ncu_step_b16.csv python bench.py --profile-step > gpurun_out/ncu_step.log 2>&1; echo "ncu step exit $?"
timeout 300 python tools/micro_post.py --out gpurun_out/micro_post_b32.json 2>&1 | tail -12
timeout 120 python tools/micro_kp.py | tail -1 | tee gpurun_out/micro_keypoints.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:keypoints_decode --launch-skip 3 -c 1 -f -o gpurun_out/ncu_keypoints_decode python tools/micro_kp.py --iters 2 > gpurun_out/ncu_keypoints_decode.log 2>&1; echo "ncu kp exit $?"
