#!/bin/bash
# parity tests of the solver kernels, A/B timings, then an ncu capture of the warp solver
OUT=gpurun_out/r02c
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py tests/test_gpu_host_api.py -m gpu -x -q > $OUT/pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest.log
tail -4 $OUT/pytest.log
B="python bench.py --no-cpu-baseline --no-gradient --steps 3 --warmup 2"
for W in nk rbc; do
  timeout 600 $B --workload $W > $OUT/bench_${W}_warp.json 2> $OUT/bench_${W}_warp.err
done
GECON_CR_KERNEL=warp timeout 600 $B --workload large > $OUT/bench_large_warp.json 2> $OUT/bench_large_warp.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02c/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d["value"]), d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["draw_outcomes"]["ok"])
    except Exception as e:
        print(f, "ERR", e)
PY
bash scripts/gpu_r02_b.sh cr_warp nk > $OUT/ncu.log 2>&1; python scripts/ncu_key_metrics.py gpurun_out/r02b/cr_warp_nk_raw.csv
head -30 gpurun_out/r02b/cr_warp_nk_lines.txt
