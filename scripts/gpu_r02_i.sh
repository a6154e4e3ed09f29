#!/bin/bash
# NP = 32 warp filter (composite model): GPU suite, then the large45 / large bench lines
OUT=gpurun_out/r02i
mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest.log
grep -E "^FAILED|^ERROR|passed|failed" $OUT/pytest.log | tail -15
B="python bench.py --no-cpu-baseline --no-gradient --no-extras --steps 3 --warmup 2"
for W in large45 large; do
  timeout 600 $B --workload $W > $OUT/bench_${W}.json 2> $OUT/bench_${W}.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02i/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["roofline"]["kernel_ms_per_step"].items()}, d["draw_outcomes"]["ok"], d["roofline"]["per_kernel"])
    except Exception as e:
        print(f, "ERR", e, open(f.replace('.json','.err')).read()[-600:])
PY
