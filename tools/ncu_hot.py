#!/usr/bin/env python
"""Per-opcode executed-instruction mix, stall-sample mix and hottest SASS lines of one kernel in an .ncu-rep.
  python tools/ncu_hot.py report.ncu-rep kernel-regex [n_units_for_normalisation]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
units = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.check_output(['ncu', '-i', rep, '--page', 'source', '--csv', '-k', 'regex:' + pat, '--launch-count', '1'],
                              stderr=subprocess.DEVNULL).decode()
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ci, si, ss = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('# Samples')
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
byop, samp, stalls = collections.Counter(), collections.Counter(), collections.Counter()
tot = tots = 0
lines = []
for r in rows[2:]:
    try:
        n, s = int(r[ci]), int(r[ss])
    except Exception:
        continue
    src = re.sub(r'^@!?U?P[0-9T]+\s+', '', r[si].strip())
    op = src.split()[0].split('.')[0] if src else '?'
    byop[op] += n; samp[op] += s; tot += n; tots += s
    for i, h in stall_cols:
        try:
            stalls[h] += int(r[i])
        except Exception:
            pass
    lines.append((s, n, r[0], r[si].strip()[:90]))
print("total warp-inst %d (%.1f per unit), samples %d" % (tot, tot / units, tots))
print("--- stall sample mix")
st = sum(stalls.values())
for h, v in stalls.most_common(10):
    print("  %-24s %5.1f%%" % (h, 100.0 * v / max(st, 1)))
print("--- opcode mix (executed, per unit, %% of inst, %% of samples)")
for op, n in byop.most_common(28):
    print("  %-10s %12d %9.1f %5.1f%% %5.1f%%" % (op, n, n / units, 100.0 * n / tot, 100.0 * samp[op] / max(tots, 1)))
print("--- hottest SASS lines by samples")
for s, n, addr, src in sorted(lines, reverse=True)[:28]:
    print("  %6d %5.2f%%  exec %9d  %s" % (s, 100.0 * s / max(tots, 1), n, src))
