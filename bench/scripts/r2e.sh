# Round-2 (second session), job 2: nvalid-aware conv epilogue, fused vocabulary head, split-K policy probes.
mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
python -m pytest tests/test_gpu_kernels.py tests/test_fused_resnet.py tests/test_fused_transformer.py -m gpu -q 2>&1 | tail -25 | tee $O/t_new.log
python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee $O/t_all.log
run() {
  name=$1; shift
  envs=""
  while [ "$1" != "--" ]; do envs="$envs $1"; shift; done
  shift
  env $envs timeout 600 python bench.py "$@" 2> $O/$name.err | tail -1 > $O/$name.json
  python - <<PY || tail -5 $O/$name.err
import json
l=json.load(open('$O/$name.json'))
print('$name', round(l['ms_per_step'],1), round(l['value']), round(l.get('e2e',{}).get('value',0)), l.get('gpu_launches'), {k: round(v,1) for k,v in l['last_round']['timings_ms'].items()})
PY
}
run c2_default -- --steps 5 --warmup 3
run c2_nostats HFL_CONV_STATS=0 -- --steps 5 --warmup 3
run c2_split HFL_CONV_KSPLIT=8 -- --steps 5 --warmup 3
run lm_fused -- --config wikitext --steps 3 --warmup 3
run lm_unfused HFL_LM_VOCAB_FUSED=0 -- --config wikitext --steps 3 --warmup 3
run c5_default -- --config cifar_stress --steps 3 --warmup 3
probe() {
  name=$1; shift
  env "$@" timeout 300 python bench/step_probe.py > $O/probe_$name.json 2> $O/probe_$name.err
  python -c "
import json; d=json.load(open('$O/probe_$name.json')); print('step_probe $name', {k: (round(v['gpu_us_per_step_isolated']), v['my_kernels']) for k,v in d.items() if k.startswith('level')}, {k: v for k, v in d.items() if k.startswith('concurrent')})" || tail -3 $O/probe_$name.err
}
probe default HFL_PDL=1
probe split HFL_CONV_KSPLIT=8
probe split_nostats HFL_CONV_KSPLIT=8 HFL_CONV_STATS=0
probe old_dispatch HFL_CONV_TC=auto HFL_CONV_BWD_TC=wgrad HFL_CONV_STATS=0
probe old_dispatch_nopdl HFL_CONV_TC=auto HFL_CONV_BWD_TC=wgrad HFL_CONV_STATS=0 HFL_PDL=0
