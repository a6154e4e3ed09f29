#!/bin/bash
# gradient sweep with DMMA products: gradient tests, gradient timing (bench default line without the other extras)
OUT=gpurun_out/r02r
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_gradient.py tests/test_gpu_kernels.py -m gpu -q --timeout=300 > $OUT/pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest.log
grep -E "^FAILED|^ERROR|passed|failed|Timeout" $OUT/pytest.log | tail -15
timeout 600 python bench.py --no-cpu-baseline --no-extras --steps 3 --warmup 2 > $OUT/bench_nk.json 2> $OUT/bench_nk.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02r/bench_nk.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["roofline"]["kernel_ms_per_step"].items()})
print(d["extras"].get("gradient"))
PY
for T in 64 128; do GECON_GRAD_THREADS=$T timeout 300 python bench.py --no-cpu-baseline --no-extras --steps 1 --warmup 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); g=d['extras']['gradient']; print('threads $T', round(g['value']), g['kernel_ms'])"; done
