#!/bin/bash
# usage: tools/quick_bench.sh <tag> [env assignments...]  -- short bench with the current library, one summary line
cd "$(dirname "$0")/.."
tag=$1; shift
env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 2>gpurun_out/qb_$tag.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tag', round(d['value'],2), 'Gb/s', round(d['ms_per_step'],3), 'ms', {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()}, 'launches', d['gpu_launches'])"
