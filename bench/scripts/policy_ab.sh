run() { echo "$@"; env "$@" python bench.py --steps 4 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('  ', round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['last_round']['timings_ms'].items()}, d['gpu_launches'])
"; }
run HFL_CONV_BWD_TC=wgrad HFL_CONV_TC=auto
run HFL_CONV_BWD_TC=off HFL_CONV_TC=auto
run HFL_CONV_BWD_TC=wgrad HFL_CONV_TC=off
run HFL_CONV_BWD_TC=on HFL_CONV_TC=auto
run HFL_CONV_BWD_TC=wgrad HFL_CONV_TC=on
