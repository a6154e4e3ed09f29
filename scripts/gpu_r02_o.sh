#!/bin/bash
# full GPU suite, then the round-2 profile capture
OUT=gpurun_out/r02o
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --timeout=300 > $OUT/pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest.log
grep -E "^FAILED|^ERROR|passed|failed|Timeout" $OUT/pytest.log | tail -15
cp gpurun_out/wide_prior_problems.json $OUT/ 2>/dev/null
bash scripts/capture_profiles.sh r02 > $OUT/capture.log 2>&1
tail -25 $OUT/capture.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/profiles_r02/r02_bench_default.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["roofline"]["kernel_ms_per_step"].items()}, {k:round(v["frac"],3) for k,v in d["roofline"]["per_kernel"].items()})
print(d["extras"].get("gradient"))
for n,o in d["extras"]["workloads"].items(): print(" ", n, round(o["value"]), round(o["ms_per_step"],2), {a:round(b,2) for a,b in o["roofline"]["kernel_ms_per_step"].items()}, {k:round(v["frac"],3) for k,v in o["roofline"]["per_kernel"].items()})
print(d["cpu_baseline"]["value"], d["cpu_baseline"]["single_core"], d["extras"]["config1_single_draw"])
PY
