#!/bin/bash
OUT=gpurun_out/r02d
mkdir -p $OUT
B="python bench.py --no-cpu-baseline --no-gradient --steps 3 --warmup 2"
for W in nk rbc; do
  timeout 600 $B --workload $W > $OUT/bench_${W}_w4.json 2> $OUT/bench_${W}_w4.err
  GECON_CW_WPC=13 timeout 600 $B --workload $W > $OUT/bench_${W}_w13.json 2> $OUT/bench_${W}_w13.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02d/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d["value"]), d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["draw_outcomes"]["ok"])
    except Exception as e:
        print(f, "ERR", e)
PY
