#!/bin/bash
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_dw_tiled.py -m gpu -q -x -s > gpurun_out/r2/pytest_dw.log 2>&1
tail -25 gpurun_out/r2/pytest_dw.log
