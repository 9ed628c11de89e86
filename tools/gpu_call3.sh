#!/bin/bash
mkdir -p gpurun_out/r2
python tools/dw_timeline.py 32 32 0 2>&1 | tee gpurun_out/r2/dw_timeline_32_0.txt
python tools/dw_timeline.py 64 64 1 2>&1 | tee gpurun_out/r2/dw_timeline_64_1.txt
