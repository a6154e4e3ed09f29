#!/bin/bash
# quick GPU round: Kalman / pipeline parity, then the bench lines of the main workloads (no CPU baseline)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py -m gpu -q -k "kalman or pipeline" 2>&1 | tail -6 > gpurun_out/quick_tests.log
tail -3 gpurun_out/quick_tests.log
for wl in nk rbc large large45; do
  python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/quick_bench_$wl.json 2> gpurun_out/quick_bench_$wl.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/quick_bench_$wl.json").read().strip().splitlines()[-1])
print("$wl", round(d["value"]), "evals/s", {k: round(v, 2) for k, v in d["roofline"]["kernel_ms_per_step"].items()}, d["draw_outcomes"]["ok"])
PY
done
python scripts/time_gradient.py > gpurun_out/grad_timing.json 2> gpurun_out/grad_timing.err; cat gpurun_out/grad_timing.json
