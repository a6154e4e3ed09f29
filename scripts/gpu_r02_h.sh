#!/bin/bash
# rank-p Joseph update in the warp filter: GPU suite, then medium NK / RBC / large NK bench lines at 16 and 20 warps per SM
OUT=gpurun_out/r02h
mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -q -x > $OUT/pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest.log
grep -E "^FAILED|^ERROR|passed|failed" $OUT/pytest.log | tail -15
B="python bench.py --no-cpu-baseline --no-gradient --no-extras --steps 3 --warmup 2"
for W in nk rbc large; do
  timeout 600 $B --workload $W > $OUT/bench_${W}_w16.json 2> $OUT/bench_${W}_w16.err
done
GECON_KW16_WPC=20 timeout 600 $B --workload nk > $OUT/bench_nk_w20.json 2> $OUT/bench_nk_w20.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02h/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["roofline"]["kernel_ms_per_step"].items()}, d["draw_outcomes"]["ok"], d.get("parity_spot_check"))
    except Exception as e:
        print(f, "ERR", e, open(f.replace('.json','.err')).read()[-600:])
PY
