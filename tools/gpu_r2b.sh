#!/bin/bash
# lean-builder tiled kernel: parity subset, timings
mkdir -p gpurun_out/r2b
timeout 300 python tools/time_tiled.py > gpurun_out/r2b/time_tiled.txt 2>&1; cat gpurun_out/r2b/time_tiled.txt
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_determinism.py tests/test_gpu_fullsize.py tests/test_gpu_bench_path_parity.py -q -x > gpurun_out/r2b/pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r2b/pytest.log
