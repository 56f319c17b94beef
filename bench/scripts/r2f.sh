# job 3: mma attention, LM probes (vocab head / attention variants), kernel census of the level-a step
mkdir -p gpurun_out/r2f
O=gpurun_out/r2f
python -m pytest tests/test_fused_transformer.py -m gpu -q 2>&1 | tail -8 | tee $O/t_lm.log
HFL_LM_ATTN_MMA=0 python -m pytest tests/test_fused_transformer.py -m gpu -q -k "attention or step_matches" 2>&1 | tail -3 | tee $O/t_lm_oldattn.log
probe() {
  name=$1; cfg=$2; shift; shift
  env "$@" timeout 400 python bench/step_probe2.py --config $cfg > $O/probe_$name.json 2> $O/probe_$name.err
  python -c "
import json; d=json.load(open('$O/probe_$name.json')); print('probe $name', {k: {a: (round(b, 1) if isinstance(b, float) else b) for a, b in v.items() if a != 'rate'} for k,v in d.items()})" || tail -3 $O/probe_$name.err
}
probe lm_default wikitext HFL_PDL=1
probe lm_novocab wikitext HFL_LM_VOCAB_FUSED=0
probe lm_oldattn wikitext HFL_LM_ATTN_MMA=0
probe lm_novocab_oldattn wikitext HFL_LM_VOCAB_FUSED=0 HFL_LM_ATTN_MMA=0
probe lm_nopdl wikitext HFL_PDL=0 HFL_LM_VOCAB_FUSED=0 HFL_LM_ATTN_MMA=0
HFL_CONV_TC=train HFL_CONV_BWD_TC=on python bench/kernel_census.py > $O/census_train.json 2> $O/census_train.err
HFL_CONV_TC=auto HFL_CONV_BWD_TC=wgrad HFL_CONV_STATS=0 python bench/kernel_census.py > $O/census_auto.json 2> $O/census_auto.err
python - <<'PY'
import json
for n in ("train", "auto"):
    d = json.load(open("gpurun_out/r2f/census_%s.json" % n))
    for k, v in d.items():
        print(n, k, v["total_kernels"], round(v["total_device_us"]))
        for name, c, t in sorted(v["by_kernel"], key=lambda r: -r[2])[:14]:
            print("   %3d %7.1f  %s" % (c, t, name[:100]))
PY
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
run lm_a_default -- --config wikitext --steps 3 --warmup 3
run lm_b_novocab HFL_LM_VOCAB_FUSED=0 -- --config wikitext --steps 3 --warmup 3
run lm_c_default -- --config wikitext --steps 3 --warmup 3
run lm_d_novocab HFL_LM_VOCAB_FUSED=0 -- --config wikitext --steps 3 --warmup 3
run c2_a_train HFL_CONV_TC=train HFL_CONV_BWD_TC=on -- --steps 5 --warmup 3
run c2_b_auto HFL_CONV_TC=auto HFL_CONV_BWD_TC=wgrad HFL_CONV_STATS=0 -- --steps 5 --warmup 3
run c2_c_train HFL_CONV_TC=train HFL_CONV_BWD_TC=on -- --steps 5 --warmup 3
run c2_d_auto HFL_CONV_TC=auto HFL_CONV_BWD_TC=wgrad HFL_CONV_STATS=0 -- --steps 5 --warmup 3
