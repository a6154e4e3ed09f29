#!/bin/bash
# per-model solver build + per-configuration filter build: identity test, pipeline / boundary tests, default bench line
mkdir -p gpurun_out/r02ad
timeout 500 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_boundary.py -x -q -m gpu --timeout=240 -k "not wide_prior_population and not full_size_population" > gpurun_out/r02ad/tests.log 2>&1
tail -5 gpurun_out/r02ad/tests.log
timeout 400 python bench.py > gpurun_out/r02ad/bench_default.json 2> gpurun_out/r02ad/bench.err
tail -c 600 gpurun_out/r02ad/bench_default.json; tail -3 gpurun_out/r02ad/bench.err
