#!/bin/bash
# full GPU suite; launch list of the default bench command; ncu --set full of the two dominant kernels (fused path)
OUT=gpurun_out/r02g
mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest.log
grep -E "^FAILED|^ERROR|passed|failed" $OUT/pytest.log | tail -15
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/r02_launches_nk.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gradient --no-extras > $OUT/launches.log 2>&1
for K in cr_warp kalman_ll_warp; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -o $OUT/r02_$K -f \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gradient --no-extras --draws 65536 > $OUT/ncu_$K.log 2>&1
  ncu -i $OUT/r02_$K.ncu-rep --page raw --csv > $OUT/r02_${K}_raw.csv 2>/dev/null
  python scripts/ncu_lines.py $OUT/r02_$K.ncu-rep 40 > $OUT/r02_${K}_lines.txt 2>/dev/null
done
ls -la $OUT
