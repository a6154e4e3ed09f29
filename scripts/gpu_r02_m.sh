#!/bin/bash
# full GPU suite (no -x), then one ncu capture of kalman_grad_kernel (opcode mix + hot lines)
OUT=gpurun_out/r02m
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --timeout=300 > $OUT/pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest.log
grep -E "^FAILED|^ERROR|passed|failed|Timeout" $OUT/pytest.log | tail -15
cp gpurun_out/wide_prior_problems.json $OUT/ 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kalman_grad -s 1 -c 1 -o $OUT/r02_kalman_grad -f \
    python scripts/time_gradient.py > $OUT/ncu_kalman_grad.log 2>&1
ncu -i $OUT/r02_kalman_grad.ncu-rep --page raw --csv > $OUT/r02_kalman_grad_raw.csv 2>/dev/null
python scripts/ncu_key_metrics.py $OUT/r02_kalman_grad_raw.csv
python scripts/ncu_lines.py $OUT/r02_kalman_grad.ncu-rep 30 > $OUT/r02_kalman_grad_lines.txt 2>/dev/null
python scripts/ncu_opmix.py $OUT/r02_kalman_grad.ncu-rep 0.3 1 > $OUT/r02_kalman_grad_opmix.txt 2>/dev/null
head -30 $OUT/r02_kalman_grad_opmix.txt
