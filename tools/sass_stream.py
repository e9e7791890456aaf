"""Compact one-char-per-instruction view of a SASS address range.  usage: sass_stream.py dump.sass kernel lo hi"""
import re, sys
txt = open(sys.argv[1]).read().split("Function : ")
body = [t for t in txt if sys.argv[2] in t.split("\n")[0]][0]
lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
keymap = {'DFMA': 'f', 'DMUL': 'm', 'DADD': 'a', 'LDS.128': 'L', 'LDS.64': 'l', 'LDS': 'l', 'STS.128': 'S', 'STS.64': 's', 'STS': 's', 'SHFL.BFLY': 'x',
          'MUFU.RSQ64H': 'Q', 'MUFU.RCP64H': 'R', 'WARPSYNC.ALL': '|', 'WARPSYNC': '|', 'NOP': '|', 'BRA': 'B', 'FSEL': '.', 'BAR.SYNC': '#',
          'LDL': 'd', 'STL': 'u', 'LDL.64': 'd', 'STL.64': 'u', 'LDL.128': 'd', 'STL.128': 'u'}
line = []
for l in body.splitlines():
    m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);', l)
    if not m: continue
    a = int(m.group(1), 16)
    if lo <= a <= hi:
        op = re.sub(r'^@!?U?P\w+\s+', '', m.group(2).strip()).split()[0]
        line.append(keymap.get(op, keymap.get(op.split('.')[0], ',')))
s = ''.join(line)
for i in range(0, len(s), 150): print(s[i:i + 150])
