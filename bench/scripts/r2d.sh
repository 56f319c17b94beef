# Round-2 (second session): PDL A/B + conv dispatch A/B + full GPU test suite.  Run with: gpurun --timeout 1500 -- bash bench/scripts/r2d.sh
mkdir -p gpurun_out/r2d
O=gpurun_out/r2d
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt
python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "conv" 2>&1 | tail -25 | tee $O/t_conv.log
python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee $O/t_all.log
HFL_PDL=0 python -m pytest tests/test_gpu_kernels.py tests/test_fused_resnet.py -m gpu -q 2>&1 | tail -5 | tee $O/t_nopdl.log
run() {  # name, env..., -- bench args
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
run c2_base HFL_PDL=0 HFL_CONV_STATS=0 -- --steps 5 --warmup 3
run c2_pdl HFL_PDL=1 HFL_CONV_STATS=0 -- --steps 5 --warmup 3
run c2_pdl_stats HFL_PDL=1 -- --steps 5 --warmup 3
run c2_pdl_bwd HFL_PDL=1 HFL_CONV_BWD_TC=on -- --steps 5 --warmup 3
run c2_pdl_bwd_train HFL_PDL=1 HFL_CONV_BWD_TC=on HFL_CONV_TC=train -- --steps 5 --warmup 3
run c2_pdl_bwd_train_nosplit HFL_PDL=1 HFL_CONV_BWD_TC=on HFL_CONV_TC=train HFL_CONV_KSPLIT=1 -- --steps 5 --warmup 3
run c2_pdl_bwd_train_nostats HFL_PDL=1 HFL_CONV_BWD_TC=on HFL_CONV_TC=train HFL_CONV_STATS=0 -- --steps 5 --warmup 3
run c2_nopdl_bwd_train HFL_PDL=0 HFL_CONV_BWD_TC=on HFL_CONV_TC=train -- --steps 5 --warmup 3
run lm_base HFL_PDL=0 -- --config wikitext --steps 3 --warmup 3
run lm_pdl HFL_PDL=1 -- --config wikitext --steps 3 --warmup 3
for m in 0 1; do
  HFL_PDL=$m HFL_CONV_BWD_TC=on HFL_CONV_TC=train timeout 300 python bench/step_probe.py > $O/step_probe_pdl$m.json 2> $O/step_probe_pdl$m.err
  python -c "
import json; d=json.load(open('$O/step_probe_pdl$m.json')); print('step_probe pdl=$m', {k: round(v['gpu_us_per_step_isolated']) for k,v in d.items() if k.startswith('level')})" || tail -3 $O/step_probe_pdl$m.err
done
HFL_PDL=0 timeout 300 python bench/step_probe.py > $O/step_probe_base.json 2> $O/step_probe_base.err
python -c "
import json; d=json.load(open('$O/step_probe_base.json')); print('step_probe base', {k: round(v['gpu_us_per_step_isolated']) for k,v in d.items() if k.startswith('level')})"
