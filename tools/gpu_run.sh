#!/bin/bash
# One parameterised command file for every GPU-box call:  gpurun --timeout T -- 'bash tools/gpu_run.sh STAGE [STAGE ...]'
# Every stage writes its artefacts under gpurun_out/ (copied into profiles/ by hand when they are to be judged) and prints
# one status line, so a call that runs several stages can be read from the tail of its output.
#
#   tests            pytest -m gpu (whole suite)            -> gpurun_out/pytest_gpu.log
#   tests:<expr>     pytest -m gpu -k '<expr>'              -> gpurun_out/pytest_gpu_k.log
#   smoke            __graft_entry__.smoke()
#   parity[:p,..]    tools/parity_report.py (800x1333, every engine or the listed precisions) -> r2_parity_fullsize.json
#   bench[:args]     bench.py [args] (+ per-layer table)    -> gpurun_out/bench.log, layers.txt   (args separated by '+')
#   reference        bench.py --impl reference --steps 2 --warmup 1
#   ncu-step[:args]  launch list (time + DRAM bytes) of ONE eager step -> gpurun_out/ncu_step.csv + summary
#   ncu-full:<regex>:<cmd>   one `--set full` capture of the first matching kernel of `python <cmd>` ('+' for spaces)
#   micro-post       tools/micro_post.py (BASELINE config 5; asserts against the oracle, then times)
#   convbench[:args] tools/conv_bench.py
#   sh:<cmd>         any other command ('+' for spaces)
mkdir -p gpurun_out
for stage in "$@"; do
  name="${stage%%:*}"; arg=""; [[ "$stage" == *:* ]] && arg="${stage#*:}"; arg="${arg//+/ }"
  case "$name" in
    tests)
      if [ -z "$arg" ]; then
        timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "[tests] exit $?"
        tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
      else
        timeout 1200 python -m pytest tests -m gpu -q --tb=short -s -k "$arg" > gpurun_out/pytest_gpu_k.log 2>&1; echo "[tests -k $arg] exit $?"
        tail -15 gpurun_out/pytest_gpu_k.log | cut -c1-400
      fi ;;
    smoke)
      timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "[smoke] exit $?"; tail -3 gpurun_out/smoke.log ;;
    parity)
      timeout 900 python tools/parity_report.py gpurun_out/r2_parity_fullsize.json ${arg//,/ } > gpurun_out/parity.log 2>&1; echo "[parity] exit $?"
      tail -4 gpurun_out/parity.log | cut -c1-900 ;;
    bench)
      tag=$(echo "$arg" | tr -c 'A-Za-z0-9' '_'); tag=${tag:-default}
      timeout 900 python bench.py --layers gpurun_out/layers_$tag.txt $arg > gpurun_out/bench_$tag.log 2>&1; echo "[bench $arg] exit $?"
      tail -1 gpurun_out/bench_$tag.log | cut -c1-1500 ;;
    reference)
      timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "[reference] exit $?"; tail -1 gpurun_out/bench_reference.log | cut -c1-400 ;;
    ncu-step)
      tag=$(echo "$arg" | tr -c 'A-Za-z0-9' '_'); tag=${tag:-default}
      timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv --log-file gpurun_out/ncu_step_$tag.csv python bench.py --profile-step $arg > gpurun_out/ncu_step_$tag.log 2>&1
      echo "[ncu-step $arg] exit $?"
      python tools/ncu_step_summary.py gpurun_out/ncu_step_$tag.csv --out gpurun_out/ncu_step_${tag}_summary.txt $NCU_SUMMARY_ARGS 2>&1 | tail -3 ;;
    ncu-full)
      regex="${arg%%:*}"; cmd="${arg#*:}"; tag=$(echo "$regex" | tr -c 'A-Za-z0-9' '_')
      timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$regex" --launch-skip ${NCU_SKIP:-2} -c 1 -f \
        -o gpurun_out/ncu_full_$tag python $cmd > gpurun_out/ncu_full_$tag.log 2>&1; echo "[ncu-full $regex] exit $?" ;;
    micro-post)
      timeout 600 python tools/micro_post.py --out gpurun_out/micro_post_b32.json $arg > gpurun_out/micro_post.log 2>&1; echo "[micro-post] exit $?"; tail -14 gpurun_out/micro_post.log | cut -c1-300 ;;
    convbench)
      timeout 600 python tools/conv_bench.py $arg > gpurun_out/convbench.txt 2>&1; echo "[convbench $arg] exit $?"; tail -40 gpurun_out/convbench.txt | cut -c1-200 ;;
    sh)
      timeout 1200 bash -c "$arg"; echo "[sh] exit $?" ;;
    *) echo "unknown stage $stage" ;;
  esac
done
