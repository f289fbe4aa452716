#!/usr/bin/env python
"""Aggregate `ncu --page source --print-source cuda,sass --csv` output: warp instructions executed and stall
samples per source file and per source line.  usage: ncu_source_hot.py file.csv [top_n]"""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; per_file = collections.Counter(); per_line = {}; samples_file = collections.Counter()
hdr = None
for row in csv.reader(open(path, newline="")):
    if not row: continue
    if row[0] == "File Path": cur = row[1].split("/")[-1]; continue
    if row[0] == "Function Name": continue
    if row[0] == "Line No": hdr = row; i_inst = hdr.index("Instructions Executed"); i_smp = hdr.index("# Samples"); i_thr = hdr.index("Thread Instructions Executed"); continue
    if row[0] == "" or hdr is None: continue
    try:
        inst = int(row[i_inst]); smp = int(row[i_smp]); thr = int(row[i_thr])
    except ValueError:
        continue
    per_file[cur] += inst; samples_file[cur] += smp
    per_line[(cur, row[0])] = (inst, smp, thr, row[1].strip()[:120])
tot = sum(per_file.values()); tots = sum(samples_file.values())
print(f"total warp instructions {tot}, stall samples {tots}")
for f, n in per_file.most_common():
    print(f"  {f:28s} inst {n:12d} {100*n/tot:5.1f}%   samples {100*samples_file[f]/max(tots,1):5.1f}%")
print("top lines by instructions executed:")
for (f, ln), (inst, smp, thr, src) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"  {100*inst/tot:5.1f}% smp {100*smp/max(tots,1):5.1f}% lanes {thr/max(inst,1):4.1f}  {f}:{ln}  {src}")
