#!/bin/bash
mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -q > gpurun_out/r2/pytest3.log 2>&1
grep -n "passed\|failed\|^FAILED" gpurun_out/r2/pytest3.log | head -40
python bench.py --steps 20 --warmup 5 > gpurun_out/r2/bench_a.json 2> gpurun_out/r2/bench_a.err
tail -c 3000 gpurun_out/r2/bench_a.json; tail -5 gpurun_out/r2/bench_a.err
