"""Summarise an `ncu --csv --metrics gpu__time_duration.sum` launch list by kernel name."""
import csv
import sys
import collections

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith('==')]
r = csv.DictReader(lines)
tot = collections.OrderedDict()
cnt = collections.Counter()
for row in r:
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = row['Kernel Name'].split('(')[0]
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    ns = v * {'ns': 1, 'us': 1e3, 'ms': 1e6, 'nsecond': 1, 'usecond': 1e3, 'msecond': 1e6}.get(unit, 1)
    tot[name] = tot.get(name, 0.0) + ns
    cnt[name] += 1
total = sum(tot.values())
print('total %.3f ms over %d launches' % (total / 1e6, sum(cnt.values())))
for name, ns in sorted(tot.items(), key=lambda kv: -kv[1])[:40]:
    print('%8.3f ms  %5.1f%%  x%-5d %s' % (ns / 1e6, 100 * ns / total, cnt[name], name[:90]))
