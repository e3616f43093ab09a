"""Aggregate `ncu --page source --print-source cuda,sass --csv` output per CUDA source line (instructions executed, samples)."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
cur_file = None
lines = []
hdr = None
for r in csv.reader(open(path)):
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        ie = hdr.index("Instructions Executed")
        ns = hdr.index("# Samples")
        continue
    if hdr and len(r) > ie and r[0].isdigit():
        try:
            lines.append((int(r[ie]), int(r[ns]), cur_file, int(r[0]), r[1].strip()))
        except ValueError:
            pass
tot = sum(l[0] for l in lines)
tots = sum(l[1] for l in lines)
print(f"total warp instructions {tot}, samples {tots}")
byfile = defaultdict(lambda: [0, 0])
for n, s, f, ln, src in lines:
    byfile[f][0] += n
    byfile[f][1] += s
for f, (n, s) in sorted(byfile.items(), key=lambda x: -x[1][0]):
    print(f"  {f:22s} instr {100 * n / tot:5.1f}%  samples {100 * s / max(tots, 1):5.1f}%")
lines.sort(reverse=True)
for n, s, f, ln, src in lines[:top]:
    print(f"{100 * n / tot:5.2f}% instr {100 * s / max(tots, 1):5.2f}% smp  {f}:{ln:<4d} {src[:105]}")
