#!/bin/bash
# full GPU test-suite + a short bench line
OUT=gpurun_out/tests
mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -q "$@" > $OUT/pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest.log
tail -40 $OUT/pytest.log
