"""Opcode histogram per loop of one kernel in a cuobjdump -sass dump.  usage: sass_loops.py dump.sass <kernel-substring>"""
import re, sys
from collections import Counter
txt = open(sys.argv[1]).read().split("Function : ")
body = [t for t in txt if sys.argv[2] in t.split("\n")[0]][0]
ins = []
for l in body.splitlines():
    m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);', l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
op = lambda t: re.sub(r'^@!?U?P\w+\s+', '', t).split()[0].split('.')[0]
print("total", len(ins), dict(Counter(op(t) for _, t in ins).most_common(14)))
idx = {a: i for i, (a, _) in enumerate(ins)}
for i, (a, t) in enumerate(ins):
    m = re.search(r'BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)', t)
    if m and int(m.group(1), 16) <= a and int(m.group(1), 16) in idx:
        b = ins[idx[int(m.group(1), 16)]:i + 1]
        if len(b) > 60:
            print(f"loop {int(m.group(1),16):#x}-{a:#x} n={len(b)}", dict(Counter(op(x[1]) for x in b).most_common(12)))
