#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics per kernel launch + stall breakdown + top stalled SASS lines."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg']
for w in want:
    if w in hdr:
        i = hdr.index(w); print("%-75s %-10s %s" % (w, units[i], [r[i][:10] for r in data]))
print()
for i, h in enumerate(hdr):
    if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio'):
        vals = [float(r[i]) for r in data]
        if max(vals) > 0.05:
            print("%-40s %s" % (h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')], ["%.2f" % v for v in vals]))
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    # ncu prints the source page of a launch more than once (one copy per source view); keep one section per launch
    sections = []
    for a, b in zip(secs[:-1], secs[1:]):
        sec = rows[a:b]
        if not sections or sec != sections[-1]:
            sections.append(sec)
    which = int(sys.argv[2])
    sec = sections[which]; h = sec[1]; body = sec[2:]
    print("\n" + " ".join(sec[0][:2]))
    ci = {x: i for i, x in enumerate(h)}
    stalls = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
    tot = sum(int(r[ci["# Samples"]]) for r in body)
    print("\nlaunch %d: %d samples, %d SASS lines" % (which, tot, len(body)))
    agg = {x: sum(int(r[ci[x]]) for r in body) for x in stalls}
    print({k[6:]: "%.1f%%" % (100.0 * v / tot) for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v > tot / 200})
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    for r in sorted(body, key=lambda r: -int(r[ci["# Samples"]]))[:n]:
        st = {x[6:]: int(r[ci[x]]) for x in stalls if int(r[ci[x]]) > 0}
        st = dict(sorted(st.items(), key=lambda x: -x[1])[:3])
        print(r[ci["# Samples"]].rjust(6), r[ci["Instructions Executed"]].rjust(9), r[1].strip()[:64].ljust(64), st)
