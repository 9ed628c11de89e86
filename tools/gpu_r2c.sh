#!/bin/bash
mkdir -p gpurun_out/r2c
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c/bench.json 2> gpurun_out/r2c/bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2c/bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d.get('ms_per_step_median'), d['value'], d['e2e'])
for k, v in d['roofline']['by_kind'].items():
    print("  %-16s %7.3f ms  n %5.1f  %7.1f GB/s" % (k, v['ms_per_step'], v['n_per_step'], v['GBps']))
PY
timeout 600 python -m pytest tests/test_gpu_heads.py -q -x > gpurun_out/r2c/pytest_heads.log 2>&1; echo "heads rc=$?"; tail -5 gpurun_out/r2c/pytest_heads.log
python tools/ncu_target.py pair > gpurun_out/r2c/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pair_dw -s 2 -c 2 -o gpurun_out/r2c/pairdw python tools/ncu_target.py pair > gpurun_out/r2c/ncu.log 2>&1; tail -2 gpurun_out/r2c/ncu.log
