"""Summaries of an .ncu-rep (run here, no GPU): key counters, stall reasons, per-line instruction counts, SASS opcode mix.

  python profiles/analyze.py gpurun_out/prof.ncu-rep <filters per launch> [min instr per filter-step to list a line]
"""
import csv
import io
import re
import subprocess
import sys
from collections import Counter, defaultdict

rep, B = sys.argv[1], int(sys.argv[2])
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 25.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
nl = len(rows) - 2
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size", "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
print("metric,unit," + ",".join(f"launch{k + 1}" for k in range(nl)))
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w},{units[i]}," + ",".join(rows[2 + k][i] for k in range(nl)))
i = hdr.index("smsp__inst_executed.sum")
print(f"# warp instructions per filter-step: {float(rows[2][i]) / B:.0f}")
st = []
for i, h in enumerate(hdr):
    if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued"):
        try:
            st.append((float(rows[2][i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
        except ValueError:
            pass
tot = sum(v for v, _ in st) or 1
print("# stall reasons: " + ", ".join(f"{h} {100 * v / tot:.1f}%" for v, h in sorted(st, reverse=True)[:8]))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, h2 = None, None
agg = defaultdict(lambda: [0, 0, ""])
ops = Counter()
seen_addr = set()
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        h2 = r
        ie, ns = h2.index("Instructions Executed"), h2.index("# Samples")
        continue
    if not h2 or len(r) <= ie:
        continue
    if r[0].isdigit():
        try:
            k = (cur, int(r[0]))
            agg[k][0] += int(r[ie]); agg[k][1] += int(r[ns]); agg[k][2] = r[1].strip()
        except ValueError:
            pass
    elif r[0] == "" and r[2].startswith("0x") and r[2] not in seen_addr:
        seen_addr.add(r[2])
        try:
            n = int(r[ie])
        except ValueError:
            continue
        op = re.sub(r"^@!?U?P\d+\s+", "", r[3].strip()).split()[0].split(".")[0]
        ops[op] += n
totl = sum(v[0] for v in agg.values())
tots = sum(v[1] for v in agg.values()) or 1
byfile = defaultdict(int)
for (f, _), v in agg.items():
    byfile[f] += v[0]
print("# instructions per filter-step by file: " + ", ".join(f"{f} {n / B / nl:.0f}" for f, n in sorted(byfile.items(), key=lambda x: -x[1])))
topn = sum(ops.values()) or 1
print("# SASS mix per filter-step: " + ", ".join(f"{k} {v / B / nl:.0f}" for k, v in ops.most_common(22)))
print(f"# source lines with >= {thr} warp instructions per filter-step (instr, % samples):")
for (f, ln), (n, s, text) in sorted(agg.items(), key=lambda kv: (kv[0][0], kv[0][1])):
    if n / B / nl >= thr:
        print(f"{n / B / nl:7.1f} {100 * s / tots:5.2f}%  {f}:{ln:<4d} {text[:100]}")
