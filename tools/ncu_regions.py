#!/usr/bin/env python
"""Sample share per SASS region (segments of N lines) of one kernel launch in an .ncu-rep."""
import csv, subprocess, io, sys
from collections import Counter
rep, which = sys.argv[1], int(sys.argv[2]); seg = int(sys.argv[3]) if len(sys.argv) > 3 else 100
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
s = secs[which]; h = rows[s + 1]; body = rows[s + 2:secs[which + 1]]
ci = {x: i for i, x in enumerate(h)}
tot = sum(int(r[ci["# Samples"]]) for r in body)
for a in range(0, len(body), seg):
    chunk = body[a:a + seg]
    sm = sum(int(r[ci["# Samples"]]) for r in chunk)
    ex = sorted(int(r[ci["Instructions Executed"]]) for r in chunk)[len(chunk) // 2]
    c = Counter()
    for r in chunk:
        t = r[1].strip()
        for m in ("BAR.SYNC", "UTMALDG", "UBLKCP", "SYNCS", "LDGSTS", "ST.E", "STG", "LDS.64", "STS.64", "IMAD.HI", "LDS.128", "LDL", "STL"):
            if m in t: c[m] += 1
    st = {x[6:]: sum(int(r[ci[x]]) for r in chunk) for x in h if x.startswith("stall_") and "Not" not in x}
    top = sorted(st.items(), key=lambda x: -x[1])[:3]
    if sm * 200 > tot or ex > 0:
        print("%5d-%5d  %5.1f%%  exec~%8d  %s  %s" % (a, a + seg, 100 * sm / tot, ex, dict(c), [(k, "%.1f" % (100 * v / tot)) for k, v in top]))
