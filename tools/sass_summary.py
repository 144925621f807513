#!/usr/bin/env python
"""SASS evidence of the product library (read on the CPU box): architectures in the fat binary, per-kernel opcode histogram of the
hot kernels, counts of the Blackwell / async-copy instructions and of atomics. Usage: python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "adfvm_b200", "csrc", "libadfvm_b200.so")
print("library:", os.path.relpath(LIB), "size", os.path.getsize(LIB))
print("cubins:", " ".join(sorted(set(re.findall(r"sm_\w+", subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout)))))
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
total = collections.Counter(); per = collections.defaultdict(collections.Counter); name = None
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and name:
        op = m.group(1)
        total[op.split(".")[0]] += 1
        per[name][op.split(".")[0]] += 1
print("kernels:", len(per))
keys = ["UBLKCP", "UBLKPF", "SYNCS", "LDGSTS", "SHFL", "DFMA", "DMUL", "DADD", "MUFU", "LDS", "STS", "LDG", "STG", "LDL", "STL", "ATOM", "ATOMS", "ATOMG", "RED", "HMMA", "UTCHMMA"]
print("whole library:", {k: total.get(k, 0) for k in keys})
assert total.get("ATOM", 0) + total.get("ATOMG", 0) + total.get("RED", 0) == 0, "float atomics found"
print("no ATOM / ATOMG / RED instructions anywhere: every scatter is a register accumulation in a fixed order")
for kname in sorted(per):
    if not re.search(r"FluxTileBody<double, 128, 288>|FluxGradTileBody<double, 128, 288>|GradAdjTileBody<double, 128, 288>|GradCellBody<double>|FluxTileBody<float, 128, 288>|FluxGradTileBody<float, 128, 288>", kname):
        continue
    c = per[kname]
    print("\n==", kname[:140])
    print("   instructions", sum(c.values()), {k: c.get(k, 0) for k in keys if c.get(k, 0)})
    print("   top:", ", ".join("%s %d" % kv for kv in c.most_common(14)))
