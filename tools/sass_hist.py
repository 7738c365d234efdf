#!/usr/bin/env python
"""Instruction histogram of address ranges of one kernel's SASS (cuobjdump -sass output).

    python tools/sass_hist.py file.sass 0xeb0-0x1b90 0x2f70-0x5350

Used to count what the hot loop of a kernel issues per iteration without a GPU."""
import collections
import re
import sys

pat = re.compile(r"^\s*/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)")
ranges = [tuple(int(x, 16) for x in a.split("-")) for a in sys.argv[2:]]
cnt = collections.Counter()
for line in open(sys.argv[1]):
    m = pat.match(line)
    if not m:
        continue
    a = int(m.group(1), 16)
    if ranges and not any(lo <= a < hi for lo, hi in ranges):
        continue
    op = m.group(2)
    if op in ("LDG", "STG", "LDS", "STS"):
        op += m.group(3)
    cnt[op] += 1
tot = sum(cnt.values())
fp = sum(v for k, v in cnt.items() if k in ("FADD", "FMUL", "FFMA"))
print("total", tot, "FADD/FMUL/FFMA", fp)
for k, v in cnt.most_common():
    print(f"{v:6d} {k}")
