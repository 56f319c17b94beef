#!/bin/bash
# job 8 (1 GPU): full suite with the split-K fixture, fused LM evaluation, row-sparse embedding optimiser, ghost-batch BN
set -u
mkdir -p gpurun_out/r2k
O=gpurun_out/r2k
python -m pytest tests/test_fused_transformer.py tests/test_gpu_kernels.py -m gpu -q -k "eval_forward or sparse_embedding or vocab_head or attention or shortcut_fold" 2>&1 | tail -12 | tee $O/t_new.log
python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee $O/t_all.log
run() {
  name=$1; shift
  envs=""
  while [ "$1" != "--" ]; do envs="$envs $1"; shift; done
  shift
  env $envs timeout 600 python bench.py "$@" 2> $O/$name.err | tail -1 > $O/$name.json
  python - <<PY || tail -5 $O/$name.err
import json
l=json.load(open('$O/$name.json'))
print('$name', round(l['ms_per_step'],1), round(l['value']), round(l.get('e2e',{}).get('value',0)), l.get('gpu_launches'), {k: round(v,1) for k,v in l['last_round']['timings_ms'].items()}, l['last_round'].get('test'))
PY
}
run lm_a_default -- --config wikitext --steps 3 --warmup 3
run lm_b_oldeval HFL_LM_EVAL_FUSED=0 -- --config wikitext --steps 3 --warmup 3
run lm_c_dense_emb HFL_LM_EMB_SPARSE=0 -- --config wikitext --steps 3 --warmup 3
run lm_d_default -- --config wikitext --steps 3 --warmup 3
run c2_a_default -- --steps 5 --warmup 3
run c2_b_ghostcluster HFL_BN_GHOST_CLUSTER=1 -- --steps 5 --warmup 3
run c2_c_default -- --steps 5 --warmup 3
run c2_d_nofold HFL_CONV_SC_FOLD=0 -- --steps 5 --warmup 3
probe() {
  name=$1; cfg=$2; shift; shift
  env "$@" timeout 400 python bench/step_probe2.py --config $cfg > $O/probe_$name.json 2> $O/probe_$name.err
  python -c "
import json; d=json.load(open('$O/probe_$name.json')); print('probe $name', {k: {a: (round(b, 1) if isinstance(b, float) else b) for a, b in v.items() if a != 'rate'} for k,v in d.items()})" || tail -3 $O/probe_$name.err
}
probe lm_default wikitext HFL_PDL=1
probe lm_dense_emb wikitext HFL_LM_EMB_SPARSE=0
