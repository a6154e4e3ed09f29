#!/bin/bash
# what the driver runs at round end: smoke(), the default bench line and the reference arm, with wall-clock times
OUT=gpurun_out/r02w
mkdir -p $OUT
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > $OUT/smoke.log 2>&1; tail -6 $OUT/smoke.log
( time timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err ) 2>&1 | grep real
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $OUT/bench_ref.json 2> $OUT/bench_ref.err ) 2>&1 | grep real
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02w/bench_default.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["roofline"]["kernel"], round(d["roofline"]["frac"],3), "traffic", d["roofline"]["traffic"], "launches", d["gpu_launches"])
r=json.loads(open("gpurun_out/r02w/bench_ref.json").read().strip().splitlines()[-1])
print("ref", round(r["value"]), r["cpu_baseline"]["kind"], r["cpu_baseline"]["cores"], r["e2e"])
PY
