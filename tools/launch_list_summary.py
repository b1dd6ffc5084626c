#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_list_summary.py in.csv [title]"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, iv, im = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
tot = collections.OrderedDict()
for r in rows[1:]:
    if r[im] != 'gpu__time_duration.sum':
        continue
    n, t = tot.get(r[ik], (0, 0.0))
    tot[r[ik]] = (n + 1, t + float(r[iv].replace(',', '')))
allt = sum(t for _, t in tot.values())
if len(sys.argv) > 2:
    print(sys.argv[2])
print('kernel, launches, total ns, share')
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print('%s, %d, %d, %.3f' % (k[:150], n, t, t / allt))
