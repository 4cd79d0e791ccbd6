"""Per-kernel launch counts and mean durations from an `ncu --metrics gpu__time_duration.sum --csv` log."""
import collections
import csv
import sys

with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith('==')]
d = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    v = v / 1e3 if unit in ('nsecond', 'ns') else v * 1e3 if unit in ('msecond', 'ms') else v
    d.setdefault(row['Kernel Name'].split('(')[0][-48:], []).append(v)
tot = sum(sum(v) for v in d.values())
for k, v in d.items():
    print("%-50s n=%5d  mean %8.2f us  share %5.1f%%" % (k, len(v), sum(v) / len(v), 100 * sum(v) / tot))
