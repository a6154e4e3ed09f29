"""Opcode mix of the hottest loop of a kernel from an ncu report (SASS source page): instructions executed per opcode, for the
instructions whose execution count is at least ``frac`` of the maximum (i.e. the inner loop).  Also lists shared-memory accesses with
excessive wavefronts (bank conflicts).  Usage: python scripts/ncu_opmix.py report.ncu-rep [frac=0.5] [units=1]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5; units = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Address")
iI, iS, iX, iW, iWi = (hdr.index(k) for k in ("Instructions Executed", "Source", "# Samples", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal"))
ins = []
for r in rows:
    if len(r) == len(hdr) and r[0].startswith("0x"):
        ins.append((int(r[iI]), r[iS].strip(), int(r[iX]), int(r[iW] or 0), int(r[iWi] or 0)))
mx = max(i[0] for i in ins)
loop = [i for i in ins if i[0] >= frac * mx]
mix = collections.Counter(); samp = collections.Counter()
for n, s, x, w, wi in loop:
    op = s.split()[0] if not s.startswith("@") else s.split()[1]
    op = op.split(".")[0]
    mix[op] += n; samp[op] += x
tot = sum(mix.values())
print(f"instructions in kernel {sum(i[0] for i in ins):,}; in the loop (count >= {frac} x max) {tot:,}; per unit {tot/units:.1f}")
for op, n in mix.most_common(30):
    print(f"  {op:10s} {n/units:10.1f}  {100*n/tot:5.1f}%  stall samples {samp[op]}")
print("shared accesses with excess wavefronts:")
for n, s, x, w, wi in sorted(ins, key=lambda t: t[3] - t[4], reverse=True)[:12]:
    if w > wi: print(f"  {s[:70]:70s} wavefronts {w/units:9.2f} ideal {wi/units:9.2f}")
