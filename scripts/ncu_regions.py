"""Instructions executed and stall samples of an ncu report aggregated over line ranges of one source file.
Usage: python scripts/ncu_regions.py report.ncu-rep file.cuh name:lo-hi [name:lo-hi ...]"""
import csv, subprocess, sys
rep, fname = sys.argv[1], sys.argv[2]
regions = [(a.split(":")[0], *map(int, a.split(":")[1].split("-"))) for a in sys.argv[3:]]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None; agg = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] not in ("", "Line No"):
        try: inst = int(r[hdr.index("Instructions Executed")]); samp = int(r[hdr.index("# Samples")])
        except ValueError: continue
        key = (cur, int(r[0])); a = agg.get(key, (0, 0)); agg[key] = (a[0] + inst, a[1] + samp)
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print(f"total warp-instructions {tot:,}; samples {ts:,}")
for name, a, b in regions:
    i = sum(v[0] for (f, l), v in agg.items() if f == fname and a <= l <= b); s = sum(v[1] for (f, l), v in agg.items() if f == fname and a <= l <= b)
    print(f"  {name:24s} inst {100*i/tot:5.1f}%  samples {100*s/ts:5.1f}%")
oth = {}
for (f, l), v in agg.items():
    if f != fname: oth[f] = oth.get(f, 0) + v[0]
print("  other files:", {k: round(100 * v / tot, 1) for k, v in oth.items()})
