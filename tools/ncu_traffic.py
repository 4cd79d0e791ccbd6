#!/usr/bin/env python
"""Writes / updates profiles/kernel_traffic.json: DRAM bytes per launch of named kernels, read from
`ncu --set full` captures (dram__bytes_read.sum + dram__bytes_write.sum). bench.py's roofline.traffic
reads this file, so the number always belongs to a committed capture of the current kernel version.
  python tools/ncu_traffic.py KEY=report.ncu-rep:kernel-regex[:capture-name] ..."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'profiles', 'kernel_traffic.json')


def to_bytes(val, unit):
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]
    return float(val.replace(',', '')) * scale


def main():
    rec = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for arg in sys.argv[1:]:
        key, rest = arg.split('=', 1)
        parts = rest.split(':')
        rep, pat = parts[0], re.compile(parts[1])
        out = subprocess.check_output(['ncu', '-i', rep, '--page', 'raw', '--csv'], stderr=subprocess.DEVNULL).decode()
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        tot, n, dur = 0.0, 0, 0.0
        for r in rows[2:]:
            if not pat.search(r[col['Kernel Name']]):
                continue
            tot += to_bytes(r[col['dram__bytes_read.sum']], units[col['dram__bytes_read.sum']])
            tot += to_bytes(r[col['dram__bytes_write.sum']], units[col['dram__bytes_write.sum']])
            dur += float(r[col['gpu__time_duration.sum']].replace(',', ''))
            n += 1
        if n == 0:
            raise SystemExit("no launch matching %s in %s" % (parts[1], rep))
        rec[key] = {"dram_bytes_per_launch": tot / n, "launches": n, "capture": parts[2] if len(parts) > 2 else os.path.basename(rep),
                    "duration_under_ncu": "%.4g %s" % (dur / n, units[col['gpu__time_duration.sum']])}
        print(key, rec[key])
    json.dump(rec, open(OUT, 'w'), indent=1, sort_keys=True)


if __name__ == '__main__':
    main()
