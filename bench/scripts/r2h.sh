#!/bin/bash
# job 5 (2 GPUs): multi-GPU test file, torchrun bench at N = 2 (both arms of ours: default; reference arm skipped here)
set -u
mkdir -p gpurun_out/r2h
O=gpurun_out/r2h
nvidia-smi -L > $O/gpus.txt
python -m pytest tests/test_multi_gpu.py -m gpu -q -s 2>&1 | tail -15 | tee $O/t_mgpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 2> $O/c2_n2.err | tail -1 > $O/c2_n2.json
python - <<'PY' || tail -5 gpurun_out/r2h/c2_n2.err
import json
l=json.load(open('gpurun_out/r2h/c2_n2.json'))
print('c2_n2', round(l['ms_per_step'],1), round(l['value']), round(l.get('e2e',{}).get('value',0)), l.get('gpu_launches'), {k: round(v,1) for k,v in l['last_round']['timings_ms'].items()}, l.get('clocks'))
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --config wikitext --steps 3 --warmup 3 2> $O/lm_n2.err | tail -1 > $O/lm_n2.json
python - <<'PY' || tail -5 gpurun_out/r2h/lm_n2.err
import json
l=json.load(open('gpurun_out/r2h/lm_n2.json'))
print('lm_n2', round(l['ms_per_step'],1), round(l['value']), round(l.get('e2e',{}).get('value',0)), l.get('gpu_launches'), {k: round(v,1) for k,v in l['last_round']['timings_ms'].items()})
PY
