#!/bin/bash
# final binary check (1 GPU): wgrad with partial image groups, full suite, smoke, headline bench
set -u
mkdir -p gpurun_out/r2q
O=gpurun_out/r2q
python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "wgrad" 2>&1 | tail -4 | tee $O/t_wgrad.log
python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $O/t_all.log
python __graft_entry__.py --smoke 2>&1 | tail -3 | tee $O/smoke.log
timeout 600 python bench.py --steps 6 --warmup 3 2> $O/c2.err | tail -1 > $O/c2.json
python - <<'PY' || tail -5 gpurun_out/r2q/c2.err
import json
l=json.load(open('gpurun_out/r2q/c2.json'))
print('c2', round(l['ms_per_step'],1), round(l['value']), round(l['e2e']['value']), l['gpu_launches'], {k: round(v,1) for k,v in l['last_round']['timings_ms'].items()}, l['clocks'])
PY
