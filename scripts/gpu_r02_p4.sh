#!/bin/bash
# two-GPU validation of the N > 1 path exactly as the driver launches it (fused pipeline, SMC-stage and strong-scaling extras)
OUT=gpurun_out/r02p4
mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 5 --warmup 3 > $OUT/bench_4gpu.json 2> $OUT/bench_4gpu.err
echo "rc=$?"
tail -3 $OUT/bench_4gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > $OUT/bench_4gpu_ref.json 2>> $OUT/bench_4gpu.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02p4/bench_4gpu.json").read().splitlines() if l.startswith("{")][-1])
print("n_gpus", d["n_gpus"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],2), d["scaling"], d["clocks"])
print({k:(v if not isinstance(v,dict) else {a:b for a,b in v.items() if a in ("value","ms_per_step","allgather_ms","ms","unit","total_draws","draws_per_gpu","exchange_ms","stage_ms")}) for k,v in d["extras"].items() if k in ("smc_stage","strong_scaling_point")})
r=[l for l in open("gpurun_out/r02p4/bench_4gpu_ref.json").read().splitlines() if l.startswith("{")]
print("ref lines", len(r), r[-1][:300] if r else None)
PY
