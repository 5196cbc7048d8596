#!/usr/bin/env python
"""Join ncu's SASS-level sampling (ncu -i X.ncu-rep --page source --csv) with nvdisasm -g line info -> hot SOURCE lines.
usage: python profiles/hotlines.py <ncu source csv> <nvdisasm -g -c dump> <kernel mangled-name substring> [N]"""
import csv, re, sys, collections
src_csv, disasm, kname = sys.argv[1:4]; N = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# 1. offset -> (file,line) from nvdisasm
lines = open(disasm, errors="replace").read().splitlines()
start = next(i for i, l in enumerate(lines) if ".text." in l and kname in l and l.strip().startswith(".section"))
cur = ("?", 0); off2line = {}; 
for l in lines[start + 1:]:
    if l.strip().startswith(".section") or l.startswith("//-----"):
        if off2line: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(\S.*?);", l)
    if m: off2line[int(m.group(1), 16)] = (cur, m.group(2))
# 2. samples per address from ncu
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; ai, si, ii = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = rows[2:]; base = int(data[0][ai], 16)
agg = collections.Counter(); ins = collections.Counter(); tot = 0; toti = 0
for r in data:
    off = int(r[ai], 16) - base; s = int(r[si]); n = int(r[ii]); tot += s; toti += n
    key = off2line.get(off, (("?", 0), ""))[0]
    agg[key] += s; ins[key] += n
print(f"total samples {tot}, warp-instructions {toti}")
for key, s in agg.most_common(N):
    print(f"{100*s/tot:6.2f}% samples {100*ins[key]/toti:6.2f}% instr  {key[0]}:{key[1]}")
