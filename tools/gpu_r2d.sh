#!/bin/bash
mkdir -p gpurun_out/r2d
timeout 300 python tools/time_dw.py > gpurun_out/r2d/time_dw.txt 2>&1; cat gpurun_out/r2d/time_dw.txt
timeout 900 python -m pytest tests/test_gpu_heads.py tests/test_gpu_ops.py tests/test_gpu_nets.py tests/test_gpu_bench_path_parity.py tests/test_gpu_fullsize.py -q -x > gpurun_out/r2d/pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r2d/pytest.log
