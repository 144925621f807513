#!/usr/bin/env python
"""Stall-sample distribution along the SASS of kernel #idx, split at barriers / mbarrier waits (phase view).
Usage: python tools/ncu_regions.py rep [kernel_index]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
tables, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == "Address":
        cur = {"hdr": row, "rows": []}; tables.append(cur)
    elif cur is not None and len(row) == len(cur["hdr"]):
        cur["rows"].append(row)
t = tables[idx]; h = t["hdr"]
ci, ce, cs = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
tot = sum(int(r[cs] or 0) for r in t["rows"])
acc_s = acc_n = 0; start = 0
print("total samples", tot)
for i, r in enumerate(t["rows"]):
    acc_s += int(r[cs] or 0); acc_n += int(r[ce] or 0)
    src = r[ci]
    if any(k in src for k in ("BAR.SYNC", "SYNCS.PHASECHK", "EXIT", "UBLKCP", "WARPSYNC")) or i == len(t["rows"]) - 1:
        if acc_s > 0.005 * tot:
            print("lines %5d-%5d  samples %6d (%4.1f%%)  warp-inst %10d   ends with: %s" % (start, i, acc_s, 100.0 * acc_s / tot, acc_n, src.strip()[:60]))
        else:
            continue
        acc_s = acc_n = 0; start = i + 1
