#!/bin/bash
mkdir -p gpurun_out/r2j
run() {
  name=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2j/$name.json 2> gpurun_out/r2j/$name.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2j/$name.json').read().strip().splitlines()[-1])
print("$name", round(d['ms_per_step'],2), round(d['value']/1e6,2), "M vox/s  exposed", d.get('comm_exposed_ms'), [round(r['ms_per_step'],2) for r in d.get('per_rank',[])])
PY
}
run default X=1
run maxctas8 NCCL_MAX_CTAS=8
run maxctas4 NCCL_MAX_CTAS=4
run maxctas2 NCCL_MAX_CTAS=2
