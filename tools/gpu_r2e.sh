#!/bin/bash
# round-2 checkpoint: GPU parity suite, bench line, smoke, ncu launch list of the same bench command
mkdir -p gpurun_out/r2e
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2e/pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2e/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2e/bench.json 2> gpurun_out/r2e/bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r2e/bench.json; tail -3 gpurun_out/r2e/bench.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2e/smoke.log 2>&1; tail -2 gpurun_out/r2e/smoke.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2e/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2e/ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r2e/launches.csv
