#!/bin/bash
# Round-2 measurement matrix on ONE B200 (gpurun --timeout 3000 -- 'bash bench/scripts/round2_matrix.sh').
# Every BASELINE.json config through bench.py (this framework), the precision-matched fp32 arm, the resident-kernel
# engine, and the reference arm where it finishes in reasonable time.  JSON lines -> gpurun_out/matrix/.
set -u
mkdir -p gpurun_out/matrix
run() {  # name timeout args...
  local name=$1 to=$2; shift 2
  timeout $to python bench.py "$@" 2> gpurun_out/matrix/$name.err | tail -1 > gpurun_out/matrix/$name.json
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    l = json.load(open("gpurun_out/matrix/%s.json" % name))
    if "unavailable" in l:
        print("%-28s unavailable: %s" % (name, l["unavailable"]))
    else:
        e2e = l.get("e2e", {}).get("value")
        print("%-28s %9.1f ms/round %10.0f samples/s  e2e %s  dtype %s  launches %s  clocks %s %s" % (
            name, l["ms_per_step"], l["value"], ("%.0f" % e2e) if e2e else "-", l.get("dtype"), l.get("gpu_launches"),
            l.get("clocks", {}).get("sm_mhz"), l.get("clocks", {}).get("reasons")))
        lr = l.get("last_round")
        if lr:
            print("%-28s   phases %s  test %s" % ("", {k: round(v, 1) for k, v in lr["timings_ms"].items()}, lr.get("test")))
except Exception as e:
    print("%-28s FAILED (%s): %s" % (name, e, open("gpurun_out/matrix/%s.err" % name).read()[-400:]))
PY
}
run ours_cifar_dynamic      600 --config cifar_dynamic --steps 5 --warmup 3
run ours_cifar_dynamic_fp32 900 --config cifar_dynamic --dtype fp32 --steps 2 --warmup 3
HFL_MEGA=1 run mega_cifar_dynamic 600 --config cifar_dynamic --steps 3 --warmup 3
run ours_cifar_stress       600 --config cifar_stress --steps 3 --warmup 3
run ours_wikitext           600 --config wikitext --steps 3 --warmup 3
run ours_cifar_homog        900 --config cifar_homog --steps 2 --warmup 3
run ours_mnist              900 --config mnist --steps 2 --warmup 3
run ref_cifar_dynamic       600 --impl reference --config cifar_dynamic --steps 1 --warmup 1
run ref_cifar_dynamic_bf16  600 --impl reference --config cifar_dynamic --ref-autocast --steps 1 --warmup 1
run ref_cifar_stress        900 --impl reference --config cifar_stress --steps 1 --warmup 1
run ref_wikitext            1200 --impl reference --config wikitext --steps 1 --warmup 1
