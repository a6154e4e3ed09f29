#!/bin/bash
# ncu capture of the solver on the RBC workload (n = 9, NP = 16)
OUT=gpurun_out/r02x
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cr_warp -s 1 -c 1 -o $OUT/r02_cr_warp_rbc -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gradient --no-extras --workload rbc > $OUT/ncu.log 2>&1
ncu -i $OUT/r02_cr_warp_rbc.ncu-rep --page raw --csv > $OUT/r02_cr_warp_rbc_raw.csv 2>/dev/null
python scripts/ncu_key_metrics.py $OUT/r02_cr_warp_rbc_raw.csv
python scripts/ncu_regions.py $OUT/r02_cr_warp_rbc.ncu-rep cr_warp.cuh load:58-76 load_compact:77-95 norm1:96-112 gj_panel:118-202 gj_update:203-241 prod:242-267 acc_helpers:268-309 power_bound:310-356 setup_loads:366-436 iter_frags_copy:437-454 iter_products:455-507 tail:508-700
