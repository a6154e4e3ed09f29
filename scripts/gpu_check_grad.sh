#!/bin/bash
# One GPU-box call: full GPU parity suite, a racecheck pass over the new gradient kernels, and a timing of the gradient path.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/gpu_tests.log
tail -5 gpurun_out/gpu_tests.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_gradient.py -q -x \
  -k "(test_kalman_grad_matches_oracle and 10-4-3 and False) or (test_policy_adjoints_match and rbc-True)" > gpurun_out/racecheck.log 2>&1
tail -8 gpurun_out/racecheck.log
python scripts/time_gradient.py > gpurun_out/grad_timing.json 2> gpurun_out/grad_timing.err
cat gpurun_out/grad_timing.json; tail -3 gpurun_out/grad_timing.err
