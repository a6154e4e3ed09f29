#!/bin/bash
# complete-sample fast path of the warp filter: kernel / pipeline tests + nk, rbc, large bench lines
OUT=gpurun_out/r02ae
mkdir -p $OUT
timeout 500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py tests/test_gpu_host_api.py -m gpu -q --timeout=240 -k "not wide_prior_population" > $OUT/pytest.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed|Timeout" $OUT/pytest.log | tail -8
B="python bench.py --no-cpu-baseline --no-gradient --no-extras --steps 5 --warmup 3"
for W in nk large; do timeout 200 $B --workload $W > $OUT/bench_$W.json 2> $OUT/bench_$W.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02ae/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d["value"]), round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["roofline"]["kernel_ms_per_step"].items()}, d["draw_outcomes"]["ok"])
    except Exception as e:
        print(f, "ERR", e, open(f.replace('.json','.err')).read()[-600:])
PY
