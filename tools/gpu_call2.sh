#!/bin/bash
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_dw_tiled.py tests/test_gpu_ops.py -m gpu -q -x > gpurun_out/r2/pytest_dw.log 2>&1
tail -12 gpurun_out/r2/pytest_dw.log
timeout 300 python tools/time_dw.py 2>&1 | tee gpurun_out/r2/time_dw.txt
