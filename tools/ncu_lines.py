"""Warp-stall samples of an ncu source-page CSV (SASS view) aggregated per SOURCE LINE, using nvdisasm -g line info of the cubin.
usage: ncu_lines.py src.csv object.o kernel_substring [top_n]"""
import csv, sys, re, subprocess, tempfile, os, glob
from collections import Counter, defaultdict
src_csv, obj, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
dis = subprocess.run(["nvdisasm", "-g", "-c"] + glob.glob(tmp + "/*.cubin"), capture_output=True, text=True).stdout
line_of = {}; cur = None; inside = False
for ln in dis.splitlines():
    if ln.startswith(".text."): inside = kname in ln; continue
    if not inside: continue
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', ln)
    if m: cur = (m.group(1), int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', ln)
    if m: line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; data = rows[2:]; col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
a0 = int(data[0][col['Address']], 16)
samples = Counter(); execd = Counter(); reasons = defaultdict(Counter); tot = 0
for r in data:
    off = int(r[col['Address']], 16) - a0
    key = line_of.get(off)
    s = int(r[col['# Samples']] or 0); tot += s
    samples[key] += s; execd[key] += int(r[col['Instructions Executed']] or 0)
    for st in stalls: reasons[key][st.replace('stall_', '')] += int(r[col[st]] or 0)
print("total samples", tot)
srcs = {}
for key, s in samples.most_common(top):
    if key is None: print("?", s); continue
    f, l = key
    if f not in srcs:
        cand = glob.glob(os.path.join(os.path.dirname(os.path.abspath(obj)), "..", f)) + glob.glob(os.path.join(os.path.dirname(os.path.abspath(obj)), f))
        srcs[f] = open(cand[0]).read().splitlines() if cand else []
    text = srcs[f][l - 1].strip()[:90] if l - 1 < len(srcs[f]) else ""
    print(f"{f}:{l:4d} {100*s/tot:5.1f}% exec {execd[key]:10d} {dict(reasons[key].most_common(3))}  | {text}")
