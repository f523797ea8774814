"""ncu launch list (`--metrics gpu__time_duration.sum --csv`) -> per-kernel count / total time / share.  usage: launch_summary.py list.csv"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors='replace')) if len(r) > 5]
hdr = next(r for r in rows if 'Kernel Name' in r)
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    if r is hdr or len(r) <= vi or r[hdr.index('Metric Name')] != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', r[ki]).replace('ssdnerf::', '').replace('void ', '')
    v = float(r[vi].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(r[ui], 1.0)
    tot[name] += v; cnt[name] += 1
total = sum(tot.values())
print(f'# {sys.argv[1]}: {sum(cnt.values())} launches, {total / 1e3:.1f} ms of kernel time (serialised, cold-cache: shares are what counts)')
for name, v in tot.most_common(40):
    print(f'{v / 1e3:10.2f} ms {100 * v / total:6.2f} %  {cnt[name]:7d} launches  {name[:90]}')
