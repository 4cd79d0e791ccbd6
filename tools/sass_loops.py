#!/usr/bin/env python
"""Static SASS census of one kernel: every backward branch (= loop) with its address range,
instruction count and opcode histogram. Usage:
  cuobjdump -sass -fun <mangled> file.o | python tools/sass_loops.py [min_instructions]"""
import re
import sys
from collections import Counter

ins = []
for line in sys.stdin:
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
addr_index = {a: i for i, (a, _) in enumerate(ins)}
min_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20


def opcode(text):
    t = text.split()
    if t[0].startswith('@'):
        t = t[1:]
    return t[0].split('.')[0]


print("total instructions", len(ins))
loops = []
for i, (a, text) in enumerate(ins):
    m = re.search(r'BRA(?:\.U)?(?:\.DIV)?\s+(?:`\(\S+\)|0x([0-9a-f]+))', text)
    if 'BRA' in text and m and m.group(1):
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr_index:
            loops.append((addr_index[tgt], i))
for lo, hi in sorted(set(loops)):
    n = hi - lo + 1
    if n < min_n:
        continue
    c = Counter(opcode(t) for _, t in ins[lo:hi + 1])
    print("loop 0x%x..0x%x  %d instr: %s" % (ins[lo][0], ins[hi][0], n,
          ' '.join('%s=%d' % kv for kv in c.most_common(14))))
