#!/usr/bin/env python
"""Executed warp instructions and stall samples per CUDA source line of one kernel in an .ncu-rep
(needs -lineinfo and --import-source on).  python tools/ncu_lines.py report.ncu-rep kernel-regex [top]"""
import collections
import csv
import io
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.check_output(['ncu', '-i', rep, '--page', 'source', '--csv', '-k', 'regex:' + pat, '--launch-count', '1',
                               '--print-source', 'cuda,sass'], stderr=subprocess.DEVNULL).decode()
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if 'Instructions Executed' in r][0]
hdr = rows[hi]
ci, ss = hdr.index('Instructions Executed'), hdr.index('# Samples')
inst, samp, text = collections.Counter(), collections.Counter(), {}
cur = None
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    if r[0].strip():
        cur = r[0].strip()
        text[cur] = r[1].strip()[:100]
    try:
        inst[cur] += int(r[ci]); samp[cur] += int(r[ss])
    except ValueError:
        pass
tot, tots = sum(inst.values()), sum(samp.values())
print("total warp instructions %d, samples %d" % (tot, tots))
for line, n in inst.most_common(top):
    print("%9d %5.1f%%  samples %5.1f%%  L%-5s %s" % (n, 100.0 * n / tot, 100.0 * samp[line] / max(tots, 1), line, text.get(line, '')))
