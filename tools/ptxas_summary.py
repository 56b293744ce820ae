#!/usr/bin/env python3
"""Summarise `nvcc -Xptxas -v` output (stdin): one line per entry function with registers, stack and spills."""
import re, subprocess, sys
txt = sys.stdin.read()
rows = []
for m in re.finditer(r"Compiling entry function '([^']+)'.*?\n(.*?)(?=ptxas info    : Compiling entry|\Z)", txt, re.S):
    name, body = m.group(1), m.group(2)
    fp = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", body)
    used = re.search(r"Used (\d+) registers", body)
    rows.append((name, used.group(1) if used else "?", *(fp.groups() if fp else ("?",) * 3)))
names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.split("\n")
for r, n in zip(rows, names):
    n = re.sub(r"\(anonymous namespace\)::|b200zk::", "", n)
    n = re.sub(r"\(.*", "", n)
    if len(sys.argv) > 1 and sys.argv[1] not in n:
        continue
    print("%-90s regs %3s stack %5s spill st/ld %5s/%5s" % (n[:90], *r[1:]))
