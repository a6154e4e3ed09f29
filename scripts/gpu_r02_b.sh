#!/bin/bash
# ncu --set full capture of the one-warp-per-draw solver (medium NK, one 65,536-draw launch) + hot source lines.
OUT=gpurun_out/r02b
mkdir -p $OUT
K=${1:-cr_warp}
W=${2:-nk}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -o $OUT/${K}_$W -f \
    python bench.py --workload $W --steps 1 --warmup 1 --no-cpu-baseline --no-gradient --draws 65536 > $OUT/ncu_${K}_$W.log 2>&1
ncu -i $OUT/${K}_$W.ncu-rep --page raw --csv > $OUT/${K}_${W}_raw.csv 2>/dev/null
python scripts/ncu_lines.py $OUT/${K}_$W.ncu-rep 70 > $OUT/${K}_${W}_lines.txt 2>/dev/null
ncu -i $OUT/${K}_$W.ncu-rep --page source --csv --print-source sass > $OUT/${K}_${W}_sass.csv 2>/dev/null
rm -f $OUT/${K}_$W.ncu-rep.tmp
ls -la $OUT
head -50 $OUT/${K}_${W}_lines.txt
