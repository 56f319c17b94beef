#!/bin/bash
# launch census of one benchmark round of the default engine (which kernels run, whose they are)
set -u
mkdir -p gpurun_out/r2n
O=gpurun_out/r2n
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none -s 60000 -c 4000 --csv --log-file $O/launches_c2.csv \
    python bench.py --steps 1 --warmup 1 > $O/ncu_c2.log 2>&1
echo "c2 rc=$? lines=$(wc -l < $O/launches_c2.csv)"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 30000 -c 3000 --csv --log-file $O/launches_lm.csv \
    python bench.py --config wikitext --steps 1 --warmup 1 > $O/ncu_lm.log 2>&1
echo "lm rc=$? lines=$(wc -l < $O/launches_lm.csv)"
python bench/launch_census.py $O/launches_c2.csv $O/launches_c2.md | head -30
python bench/launch_census.py $O/launches_lm.csv $O/launches_lm.md | head -24
