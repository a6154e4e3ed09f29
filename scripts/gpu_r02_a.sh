#!/bin/bash
# Round 2, first GPU pass of the one-warp-per-draw solver: parity tests, then A/B timings against the CTA-per-draw kernel.
OUT=gpurun_out/r02a
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest.log
tail -5 $OUT/pytest.log
B="python bench.py --no-cpu-baseline --no-gradient --steps 3 --warmup 2"
for W in nk rbc; do
  timeout 600 $B --workload $W > $OUT/bench_${W}_warp.json 2> $OUT/bench_${W}_warp.err
  GECON_CR_KERNEL=cta timeout 600 $B --workload $W > $OUT/bench_${W}_cta.json 2> $OUT/bench_${W}_cta.err
done
timeout 600 $B --workload large > $OUT/bench_large_cta.json 2> $OUT/bench_large_cta.err
GECON_CR_KERNEL=warp timeout 600 $B --workload large > $OUT/bench_large_warp.json 2> $OUT/bench_large_warp.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02a/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d["value"]), d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["draw_outcomes"]["ok"])
    except Exception as e:
        print(f, "ERR", e)
PY
