#!/bin/bash
mkdir -p gpurun_out/r2n
timeout 600 python -m pytest tests/test_gpu_deferred_dw.py tests/test_gpu_nets.py tests/test_gpu_bench_path_parity.py -q -x 2>&1 | tail -3
for v in 1 0; do
B200SCN_DEFER_DW=$v python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null > gpurun_out/r2n/n2_defer$v.json
python - <<PY
import json
d=json.loads(open("gpurun_out/r2n/n2_defer$v.json").read().strip().splitlines()[-1])
print("N=2 defer=$v", round(d["ms_per_step"],2), round(d["value"]/1e6,2), "M vox/s  identical grads:", d.get("grads_identical_across_ranks"), "exposed", round(d.get("comm_exposed_ms",0),2))
PY
done
