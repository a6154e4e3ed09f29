#!/bin/bash
# One gpurun call that refreshes everything under profiles/ for a round tag (default r01):
#   bench lines (nk / rbc / large / large45 / smc / reference arm), ncu launch list of the default bench command,
#   ncu --set full captures of the two dominant kernels and their extracted counters.
# Usage (on the GPU box, from the repo root): bash scripts/capture_profiles.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/profiles_$TAG
mkdir -p $OUT
python bench.py > $OUT/${TAG}_bench_nk.json 2> $OUT/bench_nk.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference_arm.json 2>> $OUT/bench_nk.err
python bench.py --workload rbc --no-cpu-baseline --no-gradient > $OUT/${TAG}_bench_rbc.json 2>> $OUT/bench_nk.err
python bench.py --workload large --no-cpu-baseline --no-gradient --steps 3 > $OUT/${TAG}_bench_large.json 2>> $OUT/bench_nk.err
python bench.py --workload large45 --no-cpu-baseline --no-gradient --steps 2 --warmup 3 > $OUT/${TAG}_bench_large45.json 2>> $OUT/bench_nk.err
python bench.py --workload smc --steps 5 --warmup 3 > $OUT/${TAG}_bench_smc.json 2>> $OUT/bench_nk.err
# gradient path: log-likelihood + gradient next to the plain log-likelihood, medium NK
python scripts/time_gradient.py > $OUT/${TAG}_gradient_timing.json 2>> $OUT/bench_nk.err
# launch list of the default bench command (never a bench value)
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/${TAG}_launches_nk.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gradient > $OUT/launches.log 2>&1
# full captures: one 65,536-draw launch of each dominant kernel
for K in cr_solve kalman; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -o $OUT/${TAG}_$K -f \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gradient --draws 65536 > $OUT/ncu_$K.log 2>&1
  ncu -i $OUT/${TAG}_$K.ncu-rep --page raw --csv > $OUT/${TAG}_${K}_raw.csv 2>/dev/null
  python scripts/ncu_lines.py $OUT/${TAG}_$K.ncu-rep 40 > $OUT/${TAG}_${K}_lines.txt 2>/dev/null
done
ls -la $OUT
