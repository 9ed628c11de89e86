#!/bin/bash
mkdir -p gpurun_out/r2o
timeout 900 python -m pytest tests/test_gpu_deferred_dw.py tests/test_gpu_nets.py tests/test_gpu_heads.py tests/test_reference_models.py -q -x 2>&1 | tail -3
for v in 1 0 1; do
B200SCN_DEFER_DW=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1 defer=$v', round(d['ms_per_step'],2), round(d.get('ms_per_step_median',0),2), 'e2e', round(d['e2e']['ms_per_step'],2), round(d['value']/1e6,2))"
done
