#!/bin/bash
# n = 48 solver with two tiles in the global workspace (two CTAs per SM): solver tests + composite bench
OUT=gpurun_out/r02v
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_host_api.py tests/test_gpu_pipeline.py -m gpu -q --timeout=300 > $OUT/pytest.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed|Timeout" $OUT/pytest.log | tail -12
timeout 300 python bench.py --no-cpu-baseline --no-gradient --no-extras --steps 3 --warmup 2 --workload large45 > $OUT/bench_large45.json 2> $OUT/bench_large45.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02v/bench_large45.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["roofline"]["kernel_ms_per_step"].items()}, {k:round(v["frac"],3) for k,v in d["roofline"]["per_kernel"].items()}, d["draw_outcomes"]["ok"])
PY
