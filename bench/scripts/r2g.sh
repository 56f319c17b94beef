#!/bin/bash
# job 4 (evidence): full GPU suite with the strided dgrad, ncu --set full of the round-2 kernels, compute-sanitizer runs,
# kernel census without PDL (per-kernel durations are only meaningful there), A/B of the defaults.
set -u
mkdir -p gpurun_out/r2g
O=gpurun_out/r2g
python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "dgrad" 2>&1 | tail -5 | tee $O/t_dgrad.log
python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee $O/t_all.log
N="ncu --set full --clock-control none --import-source on"
cap() {  # name, regex, skip, count, cmd...
  name=$1; rx=$2; sk=$3; ct=$4; shift; shift; shift; shift
  timeout 240 $N -k regex:$rx -s $sk -c $ct -f -o $O/prof_$name "$@" > $O/ncu_$name.log 2>&1
  echo "ncu $name rc=$?"
}
HFL_CONV_KSPLIT=8 cap conv_train conv_fprop_tcgen05 9 3 python bench/kernels_once.py convtrain
cap vocab_bwd lm_vocab_bwd 2 1 python bench/kernels_once.py vocab
cap vocab_fwd lm_vocab_fwd 2 1 python bench/kernels_once.py vocab
cap attn_mma lm_attn_.*_mma 4 2 python bench/kernels_once.py lm
cap agg2 aggregate_v2 3 1 python bench/kernels_once.py fed
cap dist2 distribute_kernel 3 1 python bench/kernels_once.py fed
cap mega mega_kernel 1 1 python bench/kernels_once.py mega 148
cap gg gemm_grouped 2 1 python -m pytest tests/test_fused_transformer.py -q -k grouped
cap mxfp8 gemm_mxfp8 2 1 python bench/mxfp8_probe.py
ls -la $O/*.ncu-rep
for tool in memcheck racecheck synccheck; do
  log=$O/sanitize_$tool.log
  timeout 420 compute-sanitizer --tool $tool --kernel-name kns=hfl --launch-timeout 300 --print-limit 20 \
      python -m pytest tests/test_gpu_kernels.py tests/test_fused_transformer.py tests/test_fused_resnet.py tests/test_mxfp8.py -q -m gpu -p no:cacheprovider \
      -k "conv_fprop or conv_dgrad or conv_wgrad or conv_bn_act or vocab_head or attention or aggregate or sgd or masked_ce or mxfp8 or mx_quantize or gn_kernels or maxpool or nvalid or grouped_qkv" > $log 2>&1
  echo "$tool rc=$? | $(grep 'ERROR SUMMARY' $log | tail -1) | $(grep -E 'passed|failed' $log | tail -1)"
done
HFL_PDL=0 HFL_CONV_KSPLIT=1 python bench/kernel_census.py > $O/census_nopdl.json 2> $O/census_nopdl.err
HFL_PDL=0 HFL_CONV_KSPLIT=8 python bench/kernel_census.py > $O/census_nopdl_split.json 2> $O/census_nopdl_split.err
python - <<'PY'
import json
for n in ("nopdl", "nopdl_split"):
    d = json.load(open("gpurun_out/r2g/census_%s.json" % n))
    for k, v in d.items():
        print(n, k, v["total_kernels"], round(v["total_device_us"]))
        for name, c, t in sorted(v["by_kernel"], key=lambda r: -r[2])[:16]:
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
run c2_default -- --steps 5 --warmup 3
probe() {
  name=$1; cfg=$2; shift; shift
  env "$@" timeout 400 python bench/step_probe2.py --config $cfg > $O/probe_$name.json 2> $O/probe_$name.err
  python -c "
import json; d=json.load(open('$O/probe_$name.json')); print('probe $name', {k: {a: (round(b, 1) if isinstance(b, float) else b) for a, b in v.items() if a != 'rate'} for k,v in d.items()})" || tail -3 $O/probe_$name.err
}
probe c2_default cifar_dynamic HFL_PDL=1
probe c2_split cifar_dynamic HFL_CONV_KSPLIT=8
