#!/bin/bash
# guarded balancing, rank-p gradient (stable), DMMA policy adjoint, full-Q gradient, posterior helpers:
# full GPU suite (short timeouts: a hung kernel must not eat the budget), wide-prior bench, gradient timing
OUT=gpurun_out/r02l
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -x --timeout=240 > $OUT/pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest.log
grep -E "^FAILED|^ERROR|passed|failed|Timeout" $OUT/pytest.log | tail -15
cp gpurun_out/wide_prior_problems.json $OUT/ 2>/dev/null
timeout 240 python bench.py --no-cpu-baseline --no-gradient --no-extras --steps 3 --warmup 2 --workload nk_wide > $OUT/bench_nk_wide.json 2> $OUT/bench_nk_wide.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02l/bench_nk_wide.json").read().strip().splitlines()[-1])
    print(round(d["value"]), round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["roofline"]["kernel_ms_per_step"].items()}, d["draw_outcomes"]["fractions"])
except Exception as e: print("ERR", e, open("gpurun_out/r02l/bench_nk_wide.err").read()[-800:])
PY
timeout 300 python scripts/time_gradient.py > $OUT/gradient_timing.json 2> $OUT/gradient_timing.err; tail -c 1500 $OUT/gradient_timing.json; tail -3 $OUT/gradient_timing.err
GECON_PA_DFMA=1 timeout 300 python scripts/time_gradient.py > $OUT/gradient_timing_dfma_pa.json 2>> $OUT/gradient_timing.err; tail -c 700 $OUT/gradient_timing_dfma_pa.json
