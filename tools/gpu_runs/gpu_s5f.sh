#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -k dwconv 2>&1 | tail -15 | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_properties.py -m gpu -q --tb=short -k "slim_dw" 2>&1 | tail -40 | cut -c1-300
