#!/bin/bash
mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -q -s > gpurun_out/r2/pytest2.log 2>&1
grep -n "^parity\|passed\|failed\|^FAILED\|Error" gpurun_out/r2/pytest2.log | head -80
python tools/cpu_overhead.py 2>&1 | head -4
