"""Aggregate warp-stall samples of an ncu source-page CSV by address range and opcode.
usage: ncu_src.py src.csv [nbuckets]"""
import csv, sys, re
from collections import Counter, defaultdict
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = 0; byop = Counter(); byop_exec = Counter(); reason = Counter()
addr0 = int(data[0][col['Address']], 16)
per = []
for r in data:
    s = int(r[col['# Samples']] or 0); ex = int(r[col['Instructions Executed']] or 0)
    src = r[col['Source']].strip()
    op = re.sub(r'^@!?U?P\w+\s+', '', src).split()[0].split('.')[0]
    byop[op] += s; byop_exec[op] += ex; tot += s
    for st in stalls: reason[st] += int(r[col[st]] or 0)
    per.append((int(r[col['Address']], 16) - addr0, src, s, ex, {st: int(r[col[st]] or 0) for st in stalls}))
print("total samples", tot)
print("by reason", {k: v for k, v in reason.most_common(8)})
print("by opcode (samples, executed)", [(k, v, byop_exec[k]) for k, v in byop.most_common(12)])
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 24
n = len(per); step = (n + nb - 1) // nb
for b in range(0, n, step):
    chunk = per[b:b + step]
    s = sum(c[2] for c in chunk); ex = sum(c[3] for c in chunk)
    rs = Counter()
    for c in chunk:
        for k, v in c[4].items(): rs[k] += v
    print(f"{chunk[0][0]:#07x}-{chunk[-1][0]:#07x} samples {s:6d} ({100*s/tot:4.1f}%) exec {ex:9d}  top {[(k.replace('stall_',''), v) for k, v in rs.most_common(3)]}")
if len(sys.argv) > 3:
    lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
    for c in per:
        if lo <= c[0] <= hi:
            top = max(c[4].items(), key=lambda kv: kv[1])
            print(f"{c[0]:#07x} {c[2]:5d} {c[3]:8d} {top[0].replace('stall_',''):>10s} {c[1][:90]}")
