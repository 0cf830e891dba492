#!/usr/bin/env python
"""Summarise an ncu launch list (gpu__time_duration.sum csv) per kernel name."""
import csv, collections, io, sys
rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
r = list(csv.DictReader(io.StringIO(''.join(rows))))
d = collections.OrderedDict()
for x in r:
    k = x['Kernel Name'][:int(sys.argv[2]) if len(sys.argv) > 2 else 60]
    v = float(x['Metric Value'].replace(',', ''))
    d.setdefault(k, [0, 0]); d[k][0] += v; d[k][1] += 1
tot = sum(v[0] for v in d.values())
for k, v in sorted(d.items(), key=lambda kv: -kv[1][0]):
    print(f"{v[0]/1e6:9.3f} ms {v[1]:4d} {100*v[0]/tot:5.1f}% {k}")
print("total %.3f ms over %d launches" % (tot / 1e6, len(r)))
