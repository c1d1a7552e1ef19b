#!/usr/bin/env python
"""Counts of the SASS mnemonics that show what the kernels are made of (cuobjdump -sass of the built library):
    python profiles/sass_grep.py > profiles/r02_sass_grep.txt
UTMALDG = TMA tensor copies, UBLKCP = TMA bulk copies, LDGSTS = cp.async, FFMA2 / FADD2 / FMUL2 = packed FP32,
IDP = DP4A / DP2A, SYNCS = mbarrier operations."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "super-resolution-system_b200", "libsrfuse.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
want = ["UTMALDG", "UBLKCP", "LDGSTS", "FFMA2", "FADD2", "FMUL2", "IDP", "VABSDIFF4", "SYNCS", "LDS", "STS", "MUFU", "HMMA", "UTC"]
counts, total, name = collections.defaultdict(collections.Counter), collections.Counter(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name).split("(")[0]
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        total[name] += 1
        for w in want:
            if m.group(1).startswith(w):
                counts[name][w] += 1
print(f"{'kernel':58s} {'instr':>6s} " + " ".join(f"{w:>9s}" for w in want))
for k in sorted(total):
    print(f"{k[:58]:58s} {total[k]:6d} " + " ".join(f"{counts[k][w]:9d}" for w in want))
