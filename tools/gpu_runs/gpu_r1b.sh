#!/bin/bash
# tests + bench (graph / eager) + CUPTI kernel timeline
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --no-cpu-baseline --layers gpurun_out/layers_b16.txt > gpurun_out/bench_graph_b16.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_graph_b16.log | cut -c1-300
timeout 600 python bench.py --no-cpu-baseline --no-graph > gpurun_out/bench_eager_b16.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_eager_b16.log | cut -c1-300
timeout 600 python tools/trace_step.py --batch 16 --steps 3 --graph 0 --out gpurun_out/trace_b16_eager.txt > gpurun_out/trace_eager.log 2>&1; echo "trace exit $?"; head -30 gpurun_out/trace_b16_eager.txt
timeout 600 python tools/trace_step.py --batch 16 --steps 3 --graph 1 --out gpurun_out/trace_b16_graph.txt > gpurun_out/trace_graph.log 2>&1; echo "trace exit $?"; head -8 gpurun_out/trace_b16_graph.txt
