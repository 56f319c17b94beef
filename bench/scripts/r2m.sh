#!/bin/bash
# job (4 GPUs): final-state check of the >= 4-rank policies (split-K, asynchronous evaluation through run_round / collect_eval)
set -u
mkdir -p gpurun_out/r2m
O=gpurun_out/r2m
nvidia-smi -L > $O/gpus.txt
run4() {
  name=$1; shift
  envs=""
  while [ "$1" != "--" ]; do envs="$envs $1"; shift; done
  shift
  env $envs timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 4 "$@" 2> $O/$name.err | tail -1 > $O/$name.json
  python - <<PY || tail -8 $O/$name.err
import json
l=json.load(open('$O/$name.json'))
print('$name', round(l['ms_per_step'],1), round(l['value']), round(l.get('e2e',{}).get('value',0)), l.get('gpu_launches'), {k: round(v,1) for k,v in l['last_round']['timings_ms'].items()}, l['config'].get('eval_overlap')[:20], l['last_round'].get('test'), l.get('clocks'))
PY
}
run4 c2_n4_default -- --steps 6 --warmup 3
run4 c2_n4_sync -- --steps 6 --warmup 3 --sync-eval
run4 lm_n4_default -- --config wikitext --steps 3 --warmup 3
