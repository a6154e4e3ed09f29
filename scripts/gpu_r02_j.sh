#!/bin/bash
# BK count with the QR fallback + rank-p gradient kernel: full GPU suite, wide-prior bench line, gradient timing
OUT=gpurun_out/r02j
mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest.log
grep -E "^FAILED|^ERROR|passed|failed" $OUT/pytest.log | tail -15
cp gpurun_out/wide_prior_problems.json $OUT/ 2>/dev/null
timeout 600 python bench.py --no-cpu-baseline --no-gradient --no-extras --steps 3 --warmup 2 --workload nk_wide > $OUT/bench_nk_wide.json 2> $OUT/bench_nk_wide.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02j/bench_nk_wide.json").read().strip().splitlines()[-1])
    print(round(d["value"]), round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["roofline"]["kernel_ms_per_step"].items()}, d["draw_outcomes"])
except Exception as e: print("ERR", e, open("gpurun_out/r02j/bench_nk_wide.err").read()[-800:])
PY
timeout 600 python scripts/time_gradient.py > $OUT/gradient_timing.json 2> $OUT/gradient_timing.err; tail -c 1500 $OUT/gradient_timing.json
