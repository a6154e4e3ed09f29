#!/bin/bash
# warp filter with T / H / d in shared memory and 20 warps per SM: tests + A/B
OUT=gpurun_out/r02aa
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py tests/test_gpu_augmentation.py -m gpu -q --timeout=300 > $OUT/pytest.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed|Timeout" $OUT/pytest.log | tail -12
B="python bench.py --no-cpu-baseline --no-gradient --no-extras --steps 5 --warmup 3"
timeout 300 $B > $OUT/bench_nk_w20ts.json 2> $OUT/bench_nk_w20ts.err
GECON_KW16_WPC=16 timeout 300 $B > $OUT/bench_nk_w16.json 2> $OUT/bench_nk_w16.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02aa/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["roofline"]["kernel_ms_per_step"].items()}, d["draw_outcomes"]["ok"])
    except Exception as e:
        print(f, "ERR", e, open(f.replace('.json','.err')).read()[-600:])
PY
