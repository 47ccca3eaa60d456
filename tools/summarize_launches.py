"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, collections, re, sys
rows = collections.OrderedDict()
with open(sys.argv[1], newline='') as f:
    lines = [l for l in f if not l.startswith('==')]
rd = csv.DictReader(lines)
total = 0.0
for r in rd:
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', r['Kernel Name'])
    name = re.sub(r'^void ', '', name)
    v = float(r['Metric Value'].replace(',', ''))
    unit = r.get('Metric Unit', 'ns')
    v *= {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'nsecond': 1e-6, 'ms': 1.0, 'msecond': 1.0}.get(unit, 1e-6)
    d = rows.setdefault(name, [0, 0.0])
    d[0] += 1; d[1] += v; total += v
print('| kernel | launches | total ms | share | avg ms |')
print('|---|---|---|---|---|')
for name, (n, t) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
    print('| %s | %d | %.3f | %.1f%% | %.3f |' % (name[:90], n, t, 100 * t / total, t / n))
print('| **total** | %d | %.3f | | |' % (sum(v[0] for v in rows.values()), total))
