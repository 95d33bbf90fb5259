#!/bin/bash
# session-4 verification: full GPU test suite, smoke, micro post bench, one bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python tools/micro_post.py --out gpurun_out/micro_post_b32.json 2>&1 | tail -12
timeout 600 python bench.py --no-cpu-baseline --layers gpurun_out/layers_b16.txt > gpurun_out/bench_b16.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_b16.log | cut -c1-600
