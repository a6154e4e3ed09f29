#!/bin/bash
# final code: remaining GPU tests (no -x), then launch list + ncu captures of the two dominant kernels
OUT=gpurun_out/r02ag
P=gpurun_out/profiles_r02
mkdir -p $OUT $P
timeout 200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py tests/test_gpu_augmentation.py tests/test_gpu_gradient.py tests/test_gpu_host_api.py tests/test_gpu_boundary.py tests/test_gpu_moments.py -m gpu -q --timeout=120 -k "not wide_prior_population and not full_size_population" > $OUT/pytest.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed|Timeout" $OUT/pytest.log | tail -8
TAG=r02
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $P/${TAG}_launches_nk.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gradient --no-extras > $P/launches.log 2>&1
for K in cr_warp kalman_ll_warp; do
  timeout 100 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -o $P/${TAG}_$K -f \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gradient --no-extras --draws 65536 > $P/ncu_$K.log 2>&1
  ncu -i $P/${TAG}_$K.ncu-rep --page raw --csv > $P/${TAG}_${K}_raw.csv 2>/dev/null
  python scripts/ncu_key_metrics.py $P/${TAG}_${K}_raw.csv > $P/${TAG}_${K}_key_metrics.txt 2>/dev/null
  python scripts/ncu_lines.py $P/${TAG}_$K.ncu-rep 40 > $P/${TAG}_${K}_lines.txt 2>/dev/null
done
python scripts/ncu_opmix.py $P/${TAG}_kalman_ll_warp.ncu-rep 0.5 13107200 > $P/${TAG}_kalman_ll_warp_opmix.txt 2>/dev/null
python scripts/ncu_opmix.py $P/${TAG}_cr_warp.ncu-rep 0.3 65536 > $P/${TAG}_cr_warp_opmix.txt 2>/dev/null
ls $P | wc -l
