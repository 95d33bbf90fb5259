#!/bin/bash
# session 5: full GPU suite + microbench + bench after the ROIAlign column-walk kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python tools/micro_post.py --old --out gpurun_out/micro_post_b32.json 2>&1 | grep roialign
timeout 900 python bench.py > gpurun_out/bench_b16.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_b16.log | cut -c1-400
