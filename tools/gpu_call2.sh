#!/bin/bash
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_point2mask_ref.py tests/test_gpu_point2mask.py tests/test_gpu_datapath.py -m gpu -q -x -s > gpurun_out/r2/pytest_p2m.log 2>&1
grep -n "point2mask ball_query\|passed\|failed\|Error" gpurun_out/r2/pytest_p2m.log | head; tail -15 gpurun_out/r2/pytest_p2m.log
