#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r02m
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s 2>&1 | tail -15 | tee gpurun_out/r02m/test_gpu_multi.txt
