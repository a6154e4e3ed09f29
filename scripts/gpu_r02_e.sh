#!/bin/bash
# full GPU suite, then fused vs kernel-by-kernel bench lines
OUT=gpurun_out/r02e
mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest.log
grep -E "^FAILED|^ERROR|passed|failed" $OUT/pytest.log | tail -15
B="python bench.py --no-cpu-baseline --no-gradient --steps 3 --warmup 2"
for W in nk rbc large; do
  timeout 600 $B --workload $W > $OUT/bench_${W}_fused.json 2> $OUT/bench_${W}_fused.err
  GECON_FUSED=0 timeout 600 $B --workload $W > $OUT/bench_${W}_plain.json 2> $OUT/bench_${W}_plain.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02e/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["roofline"]["kernel_ms_per_step"].items()}, d["draw_outcomes"]["ok"], d["gpu_launches"])
    except Exception as e:
        print(f, "ERR", e, open(f.replace('.json','.err')).read()[-600:])
PY
