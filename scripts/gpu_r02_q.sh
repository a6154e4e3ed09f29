#!/bin/bash
# thread-per-draw filter: kernel + pipeline tests, RBC bench A/B
OUT=gpurun_out/r02q
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py tests/test_gpu_augmentation.py tests/test_gpu_gradient.py tests/test_gpu_smc.py -m gpu -q --timeout=300 > $OUT/pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest.log
grep -E "^FAILED|^ERROR|passed|failed|Timeout" $OUT/pytest.log | tail -15
B="python bench.py --no-cpu-baseline --no-gradient --no-extras --steps 5 --warmup 3 --workload rbc"
timeout 300 $B > $OUT/bench_rbc_thread.json 2> $OUT/bench_rbc_thread.err
GECON_KF_THREAD=0 timeout 300 $B > $OUT/bench_rbc_warp.json 2> $OUT/bench_rbc_warp.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02q/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["roofline"]["kernel_ms_per_step"].items()}, d["draw_outcomes"]["ok"], d.get("parity_spot_check"))
    except Exception as e:
        print(f, "ERR", e, open(f.replace('.json','.err')).read()[-600:])
PY
