run() { echo "$@"; env "$@" python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('  ', round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['last_round']['timings_ms'].items()})
"; }
run HFL_STATS_CHUNK=128 HFL_TEST_CHUNK=2500
run HFL_STATS_CHUNK=512 HFL_TEST_CHUNK=5000
run HFL_STATS_CHUNK=1000 HFL_TEST_CHUNK=10000
