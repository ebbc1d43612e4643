#!/usr/bin/env python
"""Static look at a kernel's SASS: instruction mix of every loop that contains LDS/STG.
usage: tools/sass_loops.py <obj-or-so> <mangled-name-substring>"""
import re, subprocess, sys
from collections import Counter
txt = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
for f in re.split(r'\n\s*Function : ', txt)[1:]:
    name = f.split('\n')[0]
    if sys.argv[2] not in name:
        continue
    ops = []
    for l in f.split('\n'):
        m = re.search(r'/\*([0-9a-f]{4,6})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);', l)
        if m:
            ops.append((int(m.group(1), 16), m.group(3), m.group(4)))
    print(name, len(ops), "instructions")
    addr2i = {o[0]: i for i, o in enumerate(ops)}
    for i, o in enumerate(ops):
        if o[1] == 'BRA':
            m = re.search(r'0x([0-9a-f]+)', o[2])
            if m and int(m.group(1), 16) < o[0] and int(m.group(1), 16) in addr2i:
                body = ops[addr2i[int(m.group(1), 16)]:i + 1]
                c = Counter(x[1].split('.')[0] for x in body)
                if c.get('LDS', 0) + c.get('STG', 0) > 0:
                    print(' loop %#x..%#x: %d instrs' % (body[0][0], o[0], len(body)), dict(c.most_common(14)))
