#!/bin/bash
# 2 GPUs: masked-LM config with the session's new LM features, synchronous (default) and forced asynchronous evaluation
set -u
mkdir -p gpurun_out/r2o
O=gpurun_out/r2o
run2() {
  name=$1; shift
  envs=""
  while [ "$1" != "--" ]; do envs="$envs $1"; shift; done
  shift
  env $envs timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 2 "$@" 2> $O/$name.err | tail -1 > $O/$name.json
  python - <<PY || grep -E "Error|error|timed out" $O/$name.err | head -5
import json
l=json.load(open('$O/$name.json'))
print('$name', round(l['ms_per_step'],1), round(l['value']), round(l.get('e2e',{}).get('value',0)), l.get('gpu_launches'), {k: round(v,1) for k,v in l['last_round']['timings_ms'].items()}, l['last_round'].get('test'))
PY
}
run2 lm_n2_sync -- --config wikitext --steps 3 --warmup 3
run2 lm_n2_async -- --config wikitext --steps 3 --warmup 3 --async-eval
run2 lm_n2_async_oldeval HFL_LM_EVAL_FUSED=0 -- --config wikitext --steps 3 --warmup 3 --async-eval
