"""Turn the output of scripts/capture_profiles.sh (gpurun_out/profiles_<tag>/) into the tracked files under profiles/:
bench lines, launch list, per-kernel counters (<tag>_ncu_raw_metrics.json), DRAM traffic per draw (<tag>_ncu_traffic.json)
and the hot source lines.  Usage: python scripts/extract_profiles.py [tag]"""
import csv, json, shutil, sys
from pathlib import Path

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
root = Path(__file__).resolve().parent.parent
src, dst = root / "gpurun_out" / f"profiles_{tag}", root / "profiles"
WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
]
for f in src.glob(f"{tag}_bench_*.json"):
    lines = [l for l in f.read_text().splitlines() if l.startswith("{")]
    if lines:
        (dst / f.name).write_text(lines[-1] + "\n")
for f in (list(src.glob(f"{tag}_launches_nk.csv")) + list(src.glob(f"{tag}_*_lines.txt")) + list(src.glob(f"{tag}_gradient_timing.json"))
          + list(src.glob(f"{tag}_*_opmix.txt")) + list(src.glob(f"{tag}_*_key_metrics.txt"))):
    shutil.copy(f, dst / f.name)
raw, traffic = {}, {"source": "ncu --set full --clock-control none, bench.py --draws 65536 (one chunk = one launch), medium NK"}
STAGE = {"cr_solve": "cr_solve", "kalman": "kalman_ll", "cr_warp": "cr_solve", "kalman_ll_warp": "kalman_ll"}  # capture name -> bench.py stage
for k in STAGE:
    f = src / f"{tag}_{k}_raw.csv"
    if not f.exists():
        continue
    rows = list(csv.reader(f.read_text().splitlines()))
    h, units, v = rows[0], rows[1], rows[2]
    name = v[h.index("Kernel Name")].strip()
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    m = {"kernel": name}
    for i, n in enumerate(h):
        if n in WANT or ("stalled" in n and n.endswith("per_issue_active.ratio") and "not_issued" not in n):
            try:
                m[n] = float(v[i].replace(",", "")) * scale.get(units[i], 1.0)
            except ValueError:
                m[n] = v[i]
            if n == "gpu__time_duration.sum":
                m["gpu__time_duration.unit"] = units[i]
    raw[k] = m
    draws = 65536
    if "dram__bytes_read.sum" in m:
        traffic[name.split("(")[0].replace("void ", "").replace("gecon::", "")] = {
            "stage": STAGE[k], "dram_bytes_per_draw": (m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]) / draws}
(dst / f"{tag}_ncu_raw_metrics.json").write_text(json.dumps(raw, indent=1))
(dst / f"{tag}_ncu_traffic.json").write_text(json.dumps(traffic, indent=1))
print(json.dumps({k: {kk: vv for kk, vv in m.items() if "stalled" not in kk} for k, m in raw.items()}, indent=1))
