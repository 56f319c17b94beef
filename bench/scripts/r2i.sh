#!/bin/bash
# job 6 (1 GPU): BN wide/cluster threshold probe
set -u
mkdir -p gpurun_out/r2i
O=gpurun_out/r2i
probe() {
  name=$1; cfg=$2; shift; shift
  env "$@" timeout 400 python bench/step_probe2.py --config $cfg > $O/probe_$name.json 2> $O/probe_$name.err
  python -c "
import json; d=json.load(open('$O/probe_$name.json')); print('probe $name', {k: {a: (round(b, 1) if isinstance(b, float) else b) for a, b in v.items() if a != 'rate'} for k,v in d.items()})" || tail -3 $O/probe_$name.err
}
probe split cifar_dynamic HFL_CONV_KSPLIT=8
probe split_bn200k cifar_dynamic HFL_CONV_KSPLIT=8 HFL_BN_WIDE_ELEMS=200000
probe split_bn700k cifar_dynamic HFL_CONV_KSPLIT=8 HFL_BN_WIDE_ELEMS=700000
probe nosplit_bn200k cifar_dynamic HFL_BN_WIDE_ELEMS=200000
