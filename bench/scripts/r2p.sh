#!/bin/bash
# 2 GPUs: barrier kernel without early launch_dependents -- multi-GPU tests, headline + LM runs, fused asynchronous LM evaluation
set -u
mkdir -p gpurun_out/r2p
O=gpurun_out/r2p
python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -4 | tee $O/t_mgpu.log
run2() {
  name=$1; shift
  envs=""
  while [ "$1" != "--" ]; do envs="$envs $1"; shift; done
  shift
  env $envs timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 2 "$@" 2> $O/$name.err | tail -1 > $O/$name.json
  python - <<PY || grep -E "Error|error|timed out" $O/$name.err | head -5
import json
l=json.load(open('$O/$name.json'))
print('$name', round(l['ms_per_step'],1), round(l['value']), round(l.get('e2e',{}).get('value',0)), l.get('gpu_launches'), {k: round(v,1) for k,v in l['last_round']['timings_ms'].items()}, l['last_round'].get('test'))
PY
}
run2 lm_n2_async_fused HFL_LM_EVAL_FUSED_ASYNC=1 -- --config wikitext --steps 3 --warmup 3 --async-eval
run2 c2_n2_async -- --steps 5 --warmup 3 --async-eval
run2 c2_n2_default -- --steps 5 --warmup 3
