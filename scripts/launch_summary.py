"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list (per-kernel totals, shares, per-launch ms)."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ni, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[1:]:
    n = r[ni].split('(')[0]
    v = float(r[vi].replace(',', ''))
    v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(r[ui], 1.0)
    agg.setdefault(n, []).append(v)
tot = sum(sum(v) for v in agg.values())
print('| kernel | launches | total ms | share | per-launch ms |\n|---|---:|---:|---:|---|')
for n, v in sorted(agg.items(), key=lambda x: -sum(x[1])):
    print('| `%s` | %d | %.3f | %.1f%% | %s |' % (n[:70], len(v), sum(v), 100 * sum(v) / tot, ' '.join('%.2f' % x for x in v[:6])))
