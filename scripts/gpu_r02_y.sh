#!/bin/bash
OUT=gpurun_out/r02y
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -q --timeout=300 > $OUT/pytest.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed|Timeout" $OUT/pytest.log | tail -5
( time timeout 900 python bench.py --no-cpu-baseline --no-gradient > $OUT/bench_default.json 2> $OUT/bench_default.err ) 2>&1 | grep real
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02y/bench_default.json").read().strip().splitlines()[-1])
w=d["extras"]["workloads"]["nk_wide"]
print(round(w["value"]), round(w["ms_per_step"],2), w["gate_only_bk"], w["parity_spot_check"])
PY
tail -3 $OUT/bench_default.err
