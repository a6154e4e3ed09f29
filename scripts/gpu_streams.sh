#!/bin/bash
# multi-stream experiment: chunks alternate between streams so that the tail of one kernel overlaps the next chunk's kernels
mkdir -p gpurun_out
run() {
  env "$@" python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*', round(d['value']), 'evals/s', round(d['ms_per_step'],1), 'ms', 'e2e', round(d['e2e']['value']), d['draw_outcomes']['ok'])"
}
run GECON_STREAMS=1
run GECON_STREAMS=2
run GECON_STREAMS=3
run GECON_STREAMS=4
run GECON_STREAMS=2 GECON_CHUNK=32768
run GECON_STREAMS=4 GECON_CHUNK=32768
run GECON_STREAMS=4 GECON_CHUNK=16384
run GECON_STREAMS=1 GECON_CHUNK=131072
run GECON_STREAMS=2 GECON_CHUNK=131072
