"""Summarises an .ncu-rep (read here, no GPU needed) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/<tag>_rollout.ncu-rep [top_n_lines] > profiles/<tag>_rollout.txt

Section 1: launch geometry, duration, DRAM bytes, issue / occupancy / stall figures from `--page raw`.
Section 2: the source lines (needs -lineinfo + --import-source on) that own the most warp-stall samples and
executed instructions, from `--page source`.
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEEP = re.compile(
    r"^(Kernel Name|Block Size|Grid Size)$|gpu__time_duration.sum|dram__bytes_(read|write).sum$|sm__warps_active.avg.pct_of_peak|"
    r"launch__(registers_per_thread$|shared_mem_per_block_dynamic|occupancy_limit|waves|grid_size|block_size)|"
    r"smsp__issue_active.avg.pct|smsp__inst_executed.sum$|smsp__thread_inst_executed_per_inst_executed.ratio|"
    r"smsp__average_warps_issue_stalled_.*_per_issue_active|sm__throughput.avg.pct|gpu__dram_throughput.avg.pct|"
    r"l1tex__t_sector_hit_rate.pct|lts__t_sector_hit_rate.pct|smsp__warps_eligible.avg.per_cycle_active|"
    r"sm__inst_executed_pipe_(alu|fma|lsu|xu|cbu|adu).avg.pct_of_peak_sustained_active|sass__inst_executed_local|"
    r"sm__cycles_elapsed.max$|smsp__inst_executed.avg.per_cycle_active|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$|"
    r"smsp__inst_executed_op_shared|sm__sass_inst_executed_op_shared")
print(f"# {rep}")
for ri, vals in enumerate(rows[2:]):
    print(f"\n## launch {ri}")
    for h, u, v in zip(hdr, units, vals):
        if KEEP.search(h):
            print(f"{h} [{u}] = {v}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
fname, h, lines, tot_s, tot_i = "", None, [], 0, 0
stall_tot = defaultdict(int)
for r in srows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        h = r
        c_s, c_i = h.index("# Samples"), h.index("Instructions Executed")
        stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        continue
    if h is None or r[0] in ("", "Function Name") or len(r) <= c_i:
        continue
    try:
        s_, n_ = int(float(r[c_s])), int(float(r[c_i]))
    except ValueError:
        continue
    st = []
    for i, c in stall_cols:
        try:
            v = int(float(r[i]))
        except ValueError:
            v = 0
        stall_tot[c] += v
        st.append((v, c[6:]))
    st.sort(reverse=True)
    tot_s += s_
    tot_i += n_
    lines.append((s_, n_, fname, r[0], r[1].strip(), ",".join(f"{c}:{v}" for v, c in st[:3] if v)))
if not lines:
    print("\n(no source page)")
    sys.exit(0)
print(f"\n## stall samples by reason (of {tot_s})")
print(", ".join(f"{c[6:]}={100.0 * v / max(tot_s, 1):.1f}%" for c, v in sorted(stall_tot.items(), key=lambda kv: -kv[1]) if v))
byfile = defaultdict(lambda: [0, 0])
for s_, n_, f, *_ in lines:
    byfile[f][0] += s_; byfile[f][1] += n_
print("\n## by file: samples% inst%")
for f, (a, b) in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print(f"{100.0 * a / max(tot_s, 1):6.2f} {100.0 * b / max(tot_i, 1):6.2f}  {f}")
print(f"\n## hottest source lines (of {tot_s} stall samples, {tot_i} warp instructions)")
print("samples%  inst%   file:line  source   [top stalls]")
for s_, n_, f, ln, text, st in sorted(lines, reverse=True)[:topn]:
    print(f"{100.0 * s_ / max(tot_s, 1):6.2f}  {100.0 * n_ / max(tot_i, 1):6.2f}  {f}:{ln}  {text[:100]}   [{st}]")
