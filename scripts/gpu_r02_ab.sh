#!/bin/bash
# ncu capture of the per-configuration gradient kernel (medium NK: n = 10, k = 4, p = 3)
OUT=gpurun_out/r02ab
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kalman_grad_spec -s 1 -c 1 -o $OUT/r02_kalman_grad_spec -f \
    python scripts/time_gradient.py > $OUT/ncu.log 2>&1
ncu -i $OUT/r02_kalman_grad_spec.ncu-rep --page raw --csv > $OUT/r02_kalman_grad_spec_raw.csv 2>/dev/null
python scripts/ncu_key_metrics.py $OUT/r02_kalman_grad_spec_raw.csv
python scripts/ncu_lines.py $OUT/r02_kalman_grad_spec.ncu-rep 36 > $OUT/r02_kalman_grad_spec_lines.txt 2>/dev/null
python scripts/ncu_opmix.py $OUT/r02_kalman_grad_spec.ncu-rep 0.3 1 > $OUT/r02_kalman_grad_spec_opmix.txt 2>/dev/null
head -22 $OUT/r02_kalman_grad_spec_opmix.txt
