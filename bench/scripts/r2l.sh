#!/bin/bash
# final matrix (1 GPU): smoke() + every BASELINE.json config through bench.py (this framework's arm)
set -u
mkdir -p gpurun_out/r2l
O=gpurun_out/r2l
python -m pytest tests/test_fused_transformer.py -m gpu -q -k "attention" 2>&1 | tail -4 | tee $O/t_attn.log
python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee $O/t_all.log
python __graft_entry__.py --smoke 2>&1 | tail -4 | tee $O/smoke.log
run() {
  name=$1; shift
  envs=""
  while [ "$1" != "--" ]; do envs="$envs $1"; shift; done
  shift
  env $envs timeout 900 python bench.py "$@" 2> $O/$name.err | tail -1 > $O/$name.json
  python - <<PY || tail -5 $O/$name.err
import json
l=json.load(open('$O/$name.json'))
print('$name', round(l['ms_per_step'],1), round(l['value']), round(l.get('e2e',{}).get('value',0)), l.get('gpu_launches'), {k: round(v,1) for k,v in l['last_round']['timings_ms'].items()}, l.get('clocks'))
PY
}
run c2_headline -- --steps 8 --warmup 3
run c3_wikitext -- --config wikitext --steps 4 --warmup 3
run c5_stress -- --config cifar_stress --steps 3 --warmup 3
run c4_homog -- --config cifar_homog --steps 2 --warmup 3
run c1_mnist -- --config mnist --steps 2 --warmup 3
run c2_mega HFL_MEGA=1 -- --steps 2 --warmup 3
