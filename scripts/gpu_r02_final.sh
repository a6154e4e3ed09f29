#!/bin/bash
# round-2 final validation: the whole GPU test suite, smoke(), then the profile capture (bench lines, launch list, ncu --set full)
OUT=gpurun_out/r02final
mkdir -p $OUT
timeout 560 python -m pytest tests -m gpu -q --timeout=240 > $OUT/pytest.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed|Timeout" $OUT/pytest.log | tail -12
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
bash scripts/capture_profiles.sh r02 > $OUT/capture.log 2>&1
tail -25 $OUT/capture.log
