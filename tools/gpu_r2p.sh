#!/bin/bash
# checkpoint: full GPU suite, smoke, bench (deferred weight gradients on by default), ncu launch list
mkdir -p gpurun_out/r2p
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2p/pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2p/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2p/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2p/bench.json 2> gpurun_out/r2p/bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2p/bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d.get('ms_per_step_median'), d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d.get('stalled_attempt'), d['gpu_launches']/d['steps'])
print(d['roofline']['frac'], d['roofline']['step_hbm_frac'], d['cpu_baseline']['value'])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2p/bench_ref.json 2>/dev/null; echo "ref rc=$?"; tail -c 400 gpurun_out/r2p/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2p/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2p/ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r2p/launches.csv
