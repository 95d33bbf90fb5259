#!/bin/bash
# session-5 round-end style validation + profile artefacts (1 GPU)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py --layers gpurun_out/layers_b16.txt > gpurun_out/bench_b16.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_b16.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "reference exit $?"; tail -1 gpurun_out/bench_reference.log | cut -c1-300
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/ncu_step_b16.csv python bench.py --profile-step > gpurun_out/ncu_step.log 2>&1; echo "ncu step exit $?"
timeout 600 python tools/micro_post.py --old --out gpurun_out/micro_post_b32.json 2>&1 | tail -12
timeout 600 python tools/trace_step.py --batch 16 --steps 6 --graph 1 --e2e --out gpurun_out/trace_e2e.txt > gpurun_out/trace_e2e.log 2>&1; echo "trace exit $?"; head -3 gpurun_out/trace_e2e.txt
timeout 600 python tools/trace_step.py --batch 16 --steps 3 --graph 1 --out gpurun_out/trace_b16_graph.txt > gpurun_out/trace.log 2>&1; echo "trace exit $?"; head -3 gpurun_out/trace_b16_graph.txt
