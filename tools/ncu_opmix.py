#!/usr/bin/env python
"""Opcode mix + hottest SASS lines (by warp-level executed count and by stall samples) of kernel #idx in an .ncu-rep.
Usage: python tools/ncu_opmix.py rep [kernel_index=0]"""
import csv, io, subprocess, sys, collections, re
rep = sys.argv[1]; idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
# the output holds one table per kernel, separated by header rows
tables, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == "Address":
        cur = {"hdr": row, "rows": []}; tables.append(cur)
    elif cur is not None and len(row) == len(cur["hdr"]):
        cur["rows"].append(row)
t = tables[idx]
h = t["hdr"]
ci = h.index("Source"); ce = h.index("Instructions Executed")
cs = [h.index("Warp Stall Sampling (All Samples)")]
mix = collections.Counter(); tot = 0; stalls = collections.Counter()
lines = []
for r in t["rows"]:
    op = r[ci].split()[0] if r[ci] else "?"
    if op.startswith("@"):
        op = r[ci].split()[1]
    op = op.split(".")[0]
    n = int(r[ce] or 0); s = int(r[cs[0]] or 0) if cs else 0
    mix[op] += n; tot += n; stalls[op] += s
    lines.append((n, s, r[ci]))
print("total warp-inst", tot, "sass lines", len(lines))
for op, n in mix.most_common(30):
    print("  %-10s %12d %5.1f%%   stall-samples %d" % (op, n, 100.0 * n / max(tot, 1), stalls[op]))
print("-- hottest by stall samples")
for n, s, src in sorted(lines, key=lambda x: -x[1])[:25]:
    print("  %8d %6d  %s" % (n, s, src[:110]))
