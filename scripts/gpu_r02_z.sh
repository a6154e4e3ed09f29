#!/bin/bash
OUT=gpurun_out/r02z
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py tests/test_gpu_boundary.py -m gpu -q --timeout=300 -k "bk or wide or eig or solvab or gensys" > $OUT/pytest.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed|Timeout" $OUT/pytest.log | tail -5
timeout 300 python bench.py --no-cpu-baseline --no-gradient --no-extras --steps 3 --warmup 2 --workload nk_wide 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['roofline']['kernel_ms_per_step'].items()})"
