"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (originating function of the
generic launch_n / launch_reduce lambdas included).  usage: python scripts/launch_summary.py launches.csv"""
import collections
import csv
import re
import sys

with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.defaultdict(float), collections.Counter()
for row in r:
    v = float(row[iv].replace(',', ''))
    v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row[iu], 1.0)
    name = row[ik]
    m = re.search(r'dda::(\w+)(<[^>]*>)?\(.*?lambda.*?#(\d)', name)
    if m:
        name = name.split('<')[0].replace('void ', '') + ":" + m.group(1) + (m.group(2) or '') + "#" + m.group(3)
    else:
        name = re.sub(r'\(.*', '', name)[:70]
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
print("launches %d, total %.1f us" % (sum(cnt.values()), T / 1e3))
for k, v in sorted(tot.items(), key=lambda x: -x[1])[:30]:
    print("%-64s %6d %10.1f us %5.1f%%  avg %8.1f us" % (k, cnt[k], v / 1e3, 100 * v / T, v / 1e3 / cnt[k]))
