#!/bin/bash
# gradient kernel with stored update quantities: gradient + pipeline tests, timing on the bench population
OUT=gpurun_out/r02n
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_gradient.py tests/test_gpu_pipeline.py tests/test_gpu_augmentation.py -m gpu -q --timeout=300 > $OUT/pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest.log
grep -E "^FAILED|^ERROR|passed|failed|Timeout" $OUT/pytest.log | tail -15
cp gpurun_out/wide_prior_problems.json $OUT/ 2>/dev/null
timeout 600 python bench.py --no-cpu-baseline --no-extras --steps 3 --warmup 2 > $OUT/bench_nk.json 2> $OUT/bench_nk.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02n/bench_nk.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["roofline"]["kernel_ms_per_step"].items()})
print(d["extras"].get("gradient"))
PY
