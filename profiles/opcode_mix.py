#!/usr/bin/env python
"""Opcode mix and instruction-heaviest source lines of the first kernel in an .ncu-rep.
usage: python profiles/opcode_mix.py rep.ncu-rep [source-file] [top]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
srcfile = sys.argv[2] if len(sys.argv) > 2 else "esdg_cns_b200/csrc/kernels_hex.cuh"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
cs = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(cs)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
hdr = rows[hi[0]]; ci = {h: i for i, h in enumerate(hdr)}; nh = len(hdr)
end = hi[1] if len(hi) > 1 else len(rows)
cur = None; I = collections.Counter(); ops = collections.defaultdict(collections.Counter)
for r in rows[hi[0] + 1:end]:
    if not r: continue
    if r[0].strip().isdigit(): cur = int(r[0]); continue
    if len(r) >= 4 and r[2].startswith("0x"):
        off = len(r) - nh
        try: n = float(r[ci["Instructions Executed"] + off])
        except ValueError: continue
        t = r[3].strip().split()
        op = (t[1] if t and t[0].startswith("@") else (t[0] if t else "")).split(".")[0]
        I[cur] += n; ops[cur][op] += n
src = open(srcfile).read().split("\n")
tot = sum(I.values()); allops = collections.Counter()
for l in ops:
    for k, v in ops[l].items(): allops[k] += v
print(f"total warp instructions {tot/1e6:.1f} M")
print("opcode mix (%):", [(k, round(100 * v / tot, 1)) for k, v in allops.most_common(24)])
for ln, n in I.most_common(top):
    print(f"{ln:4d} {100*n/tot:5.1f}% {[(k, int(v/1e6)) for k, v in ops[ln].most_common(4)]} {src[ln-1].strip()[:72]}")
