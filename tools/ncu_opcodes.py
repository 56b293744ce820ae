#!/usr/bin/env python3
"""Per-opcode executed-instruction and stall-sample totals of one kernel from `ncu -i rep --page source --csv`.
usage: ncu_opcodes.py <source.csv> [units]   (units = divisor for the per-unit column, e.g. warp-level additions)"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops, samp = collections.Counter(), collections.Counter()
for r in rows:
    if len(r) <= ie or r is hdr:
        continue
    try:
        n, s = int(r[ie] or 0), int(r[isamp] or 0)
    except ValueError:
        continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ia].strip())
    op = m.group(2) if m else r[ia].strip()
    ops[op] += n
    samp[op] += s
tot, tots = sum(ops.values()), max(1, sum(samp.values()))
print("total warp-instructions %d, stall samples %d" % (tot, tots))
for k, v in ops.most_common(30):
    print("%-28s %12d %9.1f/unit %5.1f%%  samples %5.1f%%" % (k, v, v / units, 100 * v / tot, 100 * samp[k] / tots))
