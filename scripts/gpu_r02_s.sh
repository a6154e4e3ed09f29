#!/bin/bash
OUT=gpurun_out/r02s
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_gradient.py -m gpu -q --timeout=300 > $OUT/pytest.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed|Timeout" $OUT/pytest.log | tail -5
timeout 600 python bench.py --no-cpu-baseline --no-extras --steps 1 --warmup 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); g=d['extras']['gradient']; print('grad', round(g['value']), g['kernel_ms'])"
