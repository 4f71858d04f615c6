#!/usr/bin/env python
"""Joins an ncu source-page CSV (SASS view) with `nvdisasm --print-line-info` of the same cubin and prints executed
warp-instructions / stall samples per CUDA source line. Development aid."""
import collections
import csv
import re
import sys

sass_path, csv_path, kernel, src_path, per = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], float(sys.argv[5])
lines = open(sass_path).read().split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text." + kernel + ":")][0]
cur, seq = None, []
for l in lines[start + 1:]:
    if l.startswith("//-----") or l.startswith(".text."):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        seq.append(cur)
rows = list(csv.reader(open(csv_path)))
hdr = rows[1]
ie, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
data = [r for r in rows[2:] if len(r) == len(hdr)]
print("sass", len(seq), "ncu", len(data))
n = min(len(seq), len(data))
agg, sagg = collections.Counter(), collections.Counter()
for i in range(n):
    agg[seq[i]] += int(data[i][ie])
    sagg[seq[i]] += int(data[i][si])
tot, stot = sum(agg.values()), sum(sagg.values())
src = open(src_path).read().split("\n")
print("total warp-instr %d = %.0f per unit" % (tot, tot / per))
for key, v in agg.most_common(45):
    f, ln = key if key else ("?", 0)
    text = src[ln - 1].strip()[:80] if f == src_path.split("/")[-1] and ln <= len(src) else ""
    print("%5.1f%% ins %5.1f%% smp %7.0f  %s:%d  %s" % (100 * v / tot, 100 * sagg[key] / stot, v / per, f, ln, text))
