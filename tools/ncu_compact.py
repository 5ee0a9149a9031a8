#!/usr/bin/env python
"""Condense an `ncu --csv` launch list (one row per kernel x metric) to one line per kernel NAME: launches, mean of every metric."""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 8 and r[0].isdigit()]
acc = defaultdict(lambda: defaultdict(list))
for r in rows:
    name = r[4].split("(")[0].split("::")[-1]
    try:
        acc[name][r[-3]].append(float(r[-1].replace(",", "")))
    except ValueError:
        pass
for name, m in acc.items():
    n = max(len(v) for v in m.values())
    print(name, "launches=%d" % n, " ".join("%s=%.4g" % (k.replace("__", "_").split(".sum")[0][-34:], sum(v) / len(v)) for k, v in sorted(m.items())))
