#!/usr/bin/env python
"""SASS listing with executed counts for source lines in [lo, hi] of a file. usage: ncu_sass.py rep cubin mangled lo hi [file]"""
import csv, re, subprocess, sys
rep, cubin, kern, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
fname = sys.argv[6] if len(sys.argv) > 6 else "pe_kernels_fused2.cu"
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
lines, ops, cur, insec = [], [], None, False
for ln in dis:
    if ln.startswith("//--------------------- .text."):
        insec = kern in ln; continue
    if ln.startswith("//--------------------- "): insec = False
    if not insec: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m: lines.append(cur); ops.append(m.group(2))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; ii = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
body = [r for r in rows[2:] if len(r) > ii]
tot = sum(int(r[ii] or 0) for r in body)
for r, l, o in zip(body, lines, ops):
    if l and l[0] == fname and lo <= l[1] <= hi:
        print("%6.2f%% %5s  L%-4d %s" % (100.0 * int(r[ii] or 0) / tot, r[isamp], l[1], o))
