#!/bin/bash
# driver-format bench lines of the other BASELINE.json configurations (N=1)
mkdir -p gpurun_out/r2h
for cfg in cfg1_unet_m16_r1_s20_b1 cfg2_fcnet_m16_r1_s20_b8 cfg5_fcnet_m16_r2_res_s100_b6 cfg4_uppool_m16_r2_res_s50_b8_head; do
  timeout 600 python bench.py --config $cfg --steps 20 --warmup 5 > gpurun_out/r2h/bench_$cfg.json 2> gpurun_out/r2h/bench_$cfg.err
  echo "$cfg rc=$?"; python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2h/bench_$cfg.json').read().strip().splitlines()[-1])
    print("  ms/step", round(d['ms_per_step'], 2), "voxels/s", round(d['value']), "e2e", round(d['e2e']['value']), "cpu", round(d['cpu_baseline']['value']) if d.get('cpu_baseline') else None)
except Exception as e:
    print("  parse failed", e)
PY
  tail -2 gpurun_out/r2h/bench_$cfg.err
done
