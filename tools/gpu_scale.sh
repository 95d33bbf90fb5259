#!/bin/bash
# bench.py on N GPUs of one box, weak and strong scaling:  gpurun --gpus N -- 'bash tools/gpu_scale.sh N [extra bench args]'
N=$1; shift
mkdir -p gpurun_out
for S in weak strong; do
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 --scaling $S --no-cpu-baseline "$@" > gpurun_out/scale_n${N}_$S.log 2>&1
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --scaling $S "$@" > gpurun_out/scale_n${N}_$S.log 2>&1
  fi
  echo "n$N $S exit $?"
  tail -1 gpurun_out/scale_n${N}_$S.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read())
    print('  value %.0f img/s (%.3f ms/step, %d img/GPU)  e2e %.0f (%.3f ms)  forward_per_call %.0f  clocks %s' % (d['value'], d['ms_per_step'], d['config']['images_per_gpu'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['forward_per_call']['value'], d['clocks']))
except Exception as e:
    print('  no JSON line:', e)"
done
