#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q --tb=short -x -k "fcos_post or golden or model" 2>&1 | tail -6 | cut -c1-250
timeout 300 python tools/micro_post.py 2>&1 | grep "nms"
CM2_NMS_VARIANT=0 timeout 300 python tools/micro_post.py 2>&1 | grep "nms"
timeout 400 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['e2e']['value'])"
