#!/bin/bash
# round-2 GPU call 1: new parity tests, host-overhead profile, sanitizer logs
mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60 > gpurun_out/r2/pytest1.log
python tools/cpu_overhead.py > gpurun_out/r2/cpu_overhead.txt 2>&1
for tool in memcheck racecheck synccheck; do
  timeout 420 compute-sanitizer --tool $tool python tools/sanitize_tiled.py > gpurun_out/r2/sanitize_$tool.log 2>&1
  echo "rc=$?" >> gpurun_out/r2/sanitize_$tool.log
done
tail -30 gpurun_out/r2/pytest1.log
