#!/usr/bin/env python
"""Per-source-line instruction counts / stall samples for one kernel of an .ncu-rep.
Joins `ncu --page source --csv` (SASS rows, in order) with `nvdisasm -g` line info of the matching cubin by instruction index.
usage: tools/ncu_lines.py <rep> <cubin> <mangled-kernel-substring> [min_pct] [demangled-substring]"""
import csv
import re
import subprocess
import sys


def main():
    rep, cubin, kern = sys.argv[1:4]
    min_pct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.7
    rkern = sys.argv[5] if len(sys.argv) > 5 else ""  # substring of the demangled name in the report ("" = first kernel)
    dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
    lines, cur, insec = [], None, False
    for ln in dis:
        if ln.startswith("//--------------------- .text."):
            insec = kern in ln
            continue
        if ln.startswith("//--------------------- "):
            insec = False
        if not insec:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            lines.append(cur)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # find the block of the wanted kernel
    start = None
    for i, r in enumerate(rows):
        if r and r[0] == "Kernel Name" and (rkern in r[1]):
            start = i
            break
    hdr = rows[start + 1]
    ii, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    body = []
    for r in rows[start + 2:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) > ii:
            body.append(r)
    if len(body) != len(lines):
        print("warning: %d SASS rows in the report vs %d in the cubin" % (len(body), len(lines)))
    agg = {}
    tot_i = tot_s = 0
    for r, l in zip(body, lines):
        n, s = int(r[ii] or 0), int(r[isamp] or 0)
        tot_i += n
        tot_s += s
        a = agg.setdefault(l, [0, 0])
        a[0] += n
        a[1] += s
    src_cache = {}
    print("total warp instructions %d, stall samples %d" % (tot_i, tot_s))
    for l, (n, s) in sorted(agg.items(), key=lambda kv: (kv[0] is None, kv[0])):
        if n >= tot_i * min_pct / 100 or s >= tot_s * min_pct / 100:
            txt = ""
            if l:
                try:
                    if l[0] not in src_cache:
                        src_cache[l[0]] = open("lives_b200/csrc/" + l[0]).read().splitlines()
                    txt = src_cache[l[0]][l[1] - 1].strip()[:100]
                except Exception:
                    pass
            print("%5.1f%% inst %5.1f%% stall  %s  %s" % (100.0 * n / tot_i, 100.0 * s / max(tot_s, 1), "%s:%d" % l if l else "?", txt))


if __name__ == "__main__":
    main()
