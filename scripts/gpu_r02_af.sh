#!/bin/bash
# t_cols (zero columns of the filter's T) + state-first ordering of the filter variables: tests, then the default bench line
OUT=gpurun_out/r02af
mkdir -p $OUT gpurun_out/profiles_r02
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py tests/test_gpu_augmentation.py tests/test_gpu_gradient.py tests/test_gpu_host_api.py -m gpu -q --timeout=200 -x -k "not wide_prior_population and not full_size_population" > $OUT/pytest.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed|Timeout" $OUT/pytest.log | tail -8
timeout 300 python bench.py > $OUT/r02_bench_default.json 2> $OUT/bench_default.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02af/r02_bench_default.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["value"]), d["roofline"]["kernel_ms_per_step"])
for k,v in (d["extras"].get("workloads") or {}).items():
    print(k, round(v["value"]), v.get("roofline",{}).get("kernel_ms_per_step"))

PY
tail -3 $OUT/bench_default.err
