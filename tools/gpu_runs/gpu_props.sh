#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_properties.py -q --tb=short -x -s 2>&1 | tail -25
