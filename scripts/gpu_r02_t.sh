#!/bin/bash
# per-configuration gradient kernel: gradient tests + A/B timing
OUT=gpurun_out/r02t
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_gradient.py tests/test_gpu_augmentation.py -m gpu -q --timeout=300 > $OUT/pytest.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed|Timeout" $OUT/pytest.log | tail -5
for S in 1 0; do GECON_GRAD_SPEC=$S timeout 600 python bench.py --no-cpu-baseline --no-extras --steps 1 --warmup 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); g=d['extras']['gradient']; print('spec=$S grad', round(g['value']), {k:round(v,2) for k,v in g['kernel_ms'].items()})"; done
