#!/bin/bash
# One gpurun call that refreshes everything under profiles/ for a round tag (default r02):
#   bench lines (default line with extras / wide prior / reference arm), ncu launch list of the default bench command,
#   ncu --set full captures of the two dominant kernels and their extracted counters, opcode mix and hot lines.
# Usage (on the GPU box, from the repo root): bash scripts/capture_profiles.sh [tag]
TAG=${1:-r02}
OUT=gpurun_out/profiles_$TAG
mkdir -p $OUT
timeout 420 python bench.py > $OUT/${TAG}_bench_default.json 2> $OUT/bench_default.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference_arm.json 2>> $OUT/bench_default.err
timeout 150 python bench.py --workload nk_wide --no-cpu-baseline --no-gradient --no-extras --steps 3 > $OUT/${TAG}_bench_nk_wide.json 2>> $OUT/bench_default.err
timeout 150 python bench.py --workload smc --steps 5 --warmup 3 --no-extras > $OUT/${TAG}_bench_smc.json 2>> $OUT/bench_default.err
# launch list of the default bench command (never a bench value)
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_launches_nk.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gradient --no-extras > $OUT/launches.log 2>&1
# full captures: one 65,536-draw launch of each dominant kernel
for K in cr_warp kalman_ll_warp; do
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -o $OUT/${TAG}_$K -f \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gradient --no-extras --draws 65536 > $OUT/ncu_$K.log 2>&1
  ncu -i $OUT/${TAG}_$K.ncu-rep --page raw --csv > $OUT/${TAG}_${K}_raw.csv 2>/dev/null
  python scripts/ncu_key_metrics.py $OUT/${TAG}_${K}_raw.csv > $OUT/${TAG}_${K}_key_metrics.txt 2>/dev/null
  python scripts/ncu_lines.py $OUT/${TAG}_$K.ncu-rep 40 > $OUT/${TAG}_${K}_lines.txt 2>/dev/null
done
python scripts/ncu_opmix.py $OUT/${TAG}_kalman_ll_warp.ncu-rep 0.5 13107200 > $OUT/${TAG}_kalman_ll_warp_opmix.txt 2>/dev/null
python scripts/ncu_opmix.py $OUT/${TAG}_cr_warp.ncu-rep 0.3 65536 > $OUT/${TAG}_cr_warp_opmix.txt 2>/dev/null
ls -la $OUT
