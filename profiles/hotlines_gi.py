#!/usr/bin/env python
"""Attribute ncu SASS-level samples / executed instructions to the OUTERMOST source line of a kernel (the call site in the kernel body), using
nvdisasm -gi inline chains.  usage: python profiles/hotlines_gi.py <ncu --page source --csv of ONE kernel> <nvdisasm -gi -c dump> <mangled-name substring> [N]"""
import csv, re, sys, collections
src_csv, disasm, kname = sys.argv[1:4]; N = int(sys.argv[4]) if len(sys.argv) > 4 else 50
lines = open(disasm, errors="replace").read().splitlines()
start = next(i for i, l in enumerate(lines) if ".text." in l and kname in l and l.strip().startswith(".section"))
chain = []; off2 = {}
pending = []
for l in lines[start + 1:]:
    if l.strip().startswith(".section") or l.startswith("//-----"):
        if off2: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        pending.append((m.group(1).split("/")[-1], int(m.group(2)))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(\S.*?);", l)
    if m:
        if pending: chain = pending; pending = []
        off2[int(m.group(1), 16)] = (chain[-1] if chain else ("?", 0), chain[0] if chain else ("?", 0), m.group(2))
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; ai, si, ii, ti = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
data = rows[2:]; base = int(data[0][ai], 16)
S = collections.Counter(); I = collections.Counter(); T = collections.Counter(); C = collections.Counter(); tot = toti = tott = 0
for r in data:
    off = int(r[ai], 16) - base; s = int(r[si]); n = int(r[ii]); t = int(r[ti]); tot += s; toti += n; tott += t
    outer = off2.get(off, (("?", 0),))[0]
    S[outer] += s; I[outer] += n; T[outer] += t; C[outer] += 1
print(f"total samples {tot}, warp-instructions {toti}, thread-instructions {tott}, SASS lines {len(data)}")
print("  samples   warp-instr  thread-instr  sass  outermost line")
for key in sorted(S, key=lambda k: (k[0], k[1])):
    if S[key] * 1000 < tot and I[key] * 1000 < toti: continue
    print(f"{100*S[key]/tot:7.2f}% {100*I[key]/toti:10.2f}% {100*T[key]/tott:12.2f}% {C[key]:5d}  {key[0]}:{key[1]}")
