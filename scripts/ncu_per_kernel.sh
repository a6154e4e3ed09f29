#!/bin/bash
# One ncu pass (speed-of-light, occupancy, launch and memory sections) over every kernel of the likelihood AND gradient path:
# scripts/time_gradient.py launches all of them (32,768 medium-NK draws).  Output: gpurun_out/per_kernel/per_kernel_raw.csv
OUT=gpurun_out/per_kernel
mkdir -p $OUT
ncu --section SpeedOfLight --section Occupancy --section LaunchStats --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis \
    --clock-control none -k regex:"gecon|kalman|cr_solve|bk_count|policy" --launch-skip 28 -c 14 -o $OUT/per_kernel -f \
    python scripts/time_gradient.py > $OUT/ncu.log 2>&1
ncu -i $OUT/per_kernel.ncu-rep --page raw --csv > $OUT/per_kernel_raw.csv 2>/dev/null
tail -3 $OUT/ncu.log
