for m in 0 1; do export HFL_CUDNN_BENCHMARK=$m; export HFL_BN_SHAPE=0
  echo "HFL_CUDNN_BENCHMARK=$m"
  python bench/step_probe.py 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin)
print({k:(round(v['gpu_us_per_step_isolated'],1),v['my_kernels']) for k,v in d.items() if k.startswith('level')})
"
  python bench.py --steps 4 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['ms_per_step'], d['last_round']['timings_ms'])
"
done
