#!/bin/bash
OUT=gpurun_out/r02san
mkdir -p $OUT
timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_gradient.py -q -x -m gpu -k "kalman_grad_matches or policy_adjoints or full_shock or pipeline_gradient_matches_oracle and rbc" > $OUT/r02_sanitizer_racecheck_gradient.log 2>&1
tail -3 $OUT/r02_sanitizer_racecheck_gradient.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_kernels.py -q -x -m gpu -k "(loglik_matches_oracle and (rbc or full_nk)) or bk_certificate or cycle_reduction_parity" > $OUT/r02_sanitizer_racecheck_pipeline.log 2>&1
tail -3 $OUT/r02_sanitizer_racecheck_pipeline.log
