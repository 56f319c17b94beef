#!/bin/bash
# job 7 (8 GPUs): headline config at N = 8 -- default, async evaluation, split-K off
set -u
mkdir -p gpurun_out/r2j
O=gpurun_out/r2j
nvidia-smi -L > $O/gpus.txt
run8() {
  name=$1; shift
  envs=""
  while [ "$1" != "--" ]; do envs="$envs $1"; shift; done
  shift
  env $envs timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 "$@" 2> $O/$name.err | tail -1 > $O/$name.json
  python - <<PY || tail -8 $O/$name.err
import json
l=json.load(open('$O/$name.json'))
print('$name', round(l['ms_per_step'],1), round(l['value']), round(l.get('e2e',{}).get('value',0)), l.get('gpu_launches'), {k: round(v,1) for k,v in l['last_round']['timings_ms'].items()}, l.get('prewarm_s'), l.get('clocks'))
PY
}
run8 c2_n8_default -- --steps 6 --warmup 3
run8 c2_n8_async -- --steps 6 --warmup 3 --async-eval
run8 c2_n8_nosplit HFL_CONV_KSPLIT=1 -- --steps 6 --warmup 3
