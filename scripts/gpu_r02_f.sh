#!/bin/bash
# GPU suite (boundary + pipeline files), then the default bench command exactly as the driver runs it
OUT=gpurun_out/r02f
mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest.log
grep -E "^FAILED|^ERROR|passed|failed" $OUT/pytest.log | tail -15
( time timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err ) 2>&1 | grep real
tail -3 $OUT/bench_default.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02f/bench_default.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],2), "launches", d["gpu_launches"], "peak", d["fp64_peak_in_run"])
print("roofline", d["roofline"]["kernel"], d["roofline"]["frac"], {k:round(v,2) for k,v in d["roofline"]["kernel_ms_per_step"].items()}, {k:round(v["frac"],3) for k,v in d["roofline"]["per_kernel"].items()})
print("cpu", {k:d["cpu_baseline"][k] for k in ("value","single_core","cores","cpu_model")}, "parity", d["parity_spot_check"])
for k,v in d["extras"].items():
    if k=="workloads":
        for n,o in v.items(): print(" ", n, round(o["value"]), round(o["ms_per_step"],2), o["roofline"]["kernel"], round(o["roofline"]["frac"],3), {a:round(b,2) for a,b in o["roofline"]["kernel_ms_per_step"].items()}, o["draw_outcomes"]["fractions"], o.get("parity_spot_check"))
    else: print(" ", k, v)
PY
