#!/usr/bin/env python
"""Static SASS instruction mix of one kernel's main loop (largest backward branch), per opcode.
usage: sass_loop_mix.py sass.txt <substring of the mangled kernel name>"""
import re, sys
from collections import Counter
txt = open(sys.argv[1]).read()
parts = re.split(r'\n\s*Function : ', txt)
sel = [p for p in parts[1:] if sys.argv[2] in p.split('\n')[0]]
assert sel, "kernel not found"
body = sel[0]
ins = []
for l in body.splitlines():
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)', l)
    if m:
        ins.append((int(m.group(1), 16), m.group(3), l))
loops = []
for a, op, l in ins:
    if op.startswith('BRA'):
        t = re.search(r'0x([0-9a-f]+)', l.split('BRA')[1])
        if t and int(t.group(1), 16) < a:
            loops.append((a - int(t.group(1), 16), int(t.group(1), 16), a))
loops.sort(reverse=True)
size, lo, hi = loops[0]
inl = [(a, op) for a, op, _ in ins if lo <= a <= hi]
c = Counter(op.split('.')[0] for _, op in inl)
print("kernel %s: %d instr total, main loop %d instr (0x%x..0x%x)" % (body.split('\n')[0][:60], len(ins), len(inl), lo, hi))
fp64 = sum(v for k, v in c.items() if k in ('DFMA', 'DMUL', 'DADD', 'DSETP', 'MUFU'))
print("  FP64-pipe %d, other %d" % (fp64, len(inl) - fp64))
print("  " + ", ".join("%s %d" % kv for kv in c.most_common(24)))
