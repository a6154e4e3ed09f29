"""Aggregate an ncu source page (``ncu -i X.ncu-rep --page source --csv --print-source cuda,sass``) per CUDA source line:
instructions executed and stall samples.  Usage: python scripts/ncu_lines.py report.ncu-rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None; hdr = None; agg = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] not in ("", "Line No"):
        try:
            inst = int(r[hdr.index("Instructions Executed")]); samp = int(r[hdr.index("# Samples")])
        except ValueError:
            continue
        agg.append((inst, samp, cur_file, r[0], r[1].strip()[:110]))
tot_i = sum(a[0] for a in agg); tot_s = sum(a[1] for a in agg)
print(f"total warp-instructions {tot_i:,}  samples {tot_s:,}")
for inst, samp, f, ln, src in sorted(agg, reverse=True)[:top]:
    print(f"{100*inst/tot_i:5.1f}% inst {100*samp/max(tot_s,1):5.1f}% stall  {f}:{ln}  {src}")
