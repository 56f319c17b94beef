#!/bin/bash
# deeper TMA ring of the conv kernel (6-9 stages) vs the 4-stage build (ops/_hfl_C_s4.so), same box: conv tests, step probes, round
set -u
mkdir -p gpurun_out/r2s
O=gpurun_out/r2s
PKG=$(ls -d heterofl*_b200)
python -m pytest tests/test_gpu_kernels.py tests/test_fused_resnet.py -m gpu -q -k "conv or fused" 2>&1 | tail -3 | tee $O/t_conv.log
probe() {
  name=$1; cfg=$2; shift; shift
  env "$@" timeout 300 python bench/step_probe2.py --config $cfg > $O/probe_$name.json 2> $O/probe_$name.err
  python -c "
import json; d=json.load(open('$O/probe_$name.json')); print('probe $name', {k: {a: (round(b, 1) if isinstance(b, float) else b) for a, b in v.items() if a != 'rate'} for k,v in d.items()})" || tail -3 $O/probe_$name.err
}
run() {
  name=$1; shift
  timeout 400 python bench.py "$@" 2> $O/$name.err | tail -1 > $O/$name.json
  python - <<PY || tail -5 $O/$name.err
import json
l=json.load(open('$O/$name.json'))
print('$name', round(l['ms_per_step'],1), round(l['value']), round(l['e2e']['value']), l['gpu_launches'], {k: round(v,1) for k,v in l['last_round']['timings_ms'].items()})
PY
}
probe deep cifar_dynamic HFL_PDL=1
probe deep_split cifar_dynamic HFL_CONV_KSPLIT=8
run c2_deep --steps 5 --warmup 3
cp $PKG/ops/_hfl_C.so /tmp/_hfl_C_deep.so
cp $PKG/ops/_hfl_C_s4.so $PKG/ops/_hfl_C.so
probe s4 cifar_dynamic HFL_PDL=1
probe s4_split cifar_dynamic HFL_CONV_KSPLIT=8
run c2_s4 --steps 5 --warmup 3
cp /tmp/_hfl_C_deep.so $PKG/ops/_hfl_C.so
