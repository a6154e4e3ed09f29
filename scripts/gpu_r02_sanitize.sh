#!/bin/bash
# compute-sanitizer over the kernels that are new or changed in round 2 (small cases only: the tools slow kernels down 10-100x)
OUT=gpurun_out/r02san
mkdir -p $OUT
K1="synthetic_sizes or bk_count or cycle_reduction_parity or intercept_meets"
timeout 700 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_kernels.py -q -x -m gpu -k "$K1" > $OUT/r02_sanitizer_memcheck_kernels.log 2>&1
tail -4 $OUT/r02_sanitizer_memcheck_kernels.log
timeout 700 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_kernels.py -q -x -m gpu -k "synthetic_sizes or bk_count" > $OUT/r02_sanitizer_racecheck_kernels.log 2>&1
tail -4 $OUT/r02_sanitizer_racecheck_kernels.log
timeout 700 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_gradient.py -q -x -m gpu -k "kalman_grad_matches or policy_adjoints or full_shock or pipeline_gradient_matches_oracle and rbc" > $OUT/r02_sanitizer_racecheck_gradient.log 2>&1
tail -4 $OUT/r02_sanitizer_racecheck_gradient.log
timeout 700 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_boundary.py tests/test_gpu_pipeline.py -q -x -m gpu -k "eig or fused or loglik_matches_oracle and rbc" > $OUT/r02_sanitizer_memcheck_pipeline.log 2>&1
tail -4 $OUT/r02_sanitizer_memcheck_pipeline.log
