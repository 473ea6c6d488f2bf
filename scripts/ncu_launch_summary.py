"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python scripts/ncu_launch_summary.py launches.csv [skip_first_n] > summary.txt"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1 + skip:]:
    v = float(r[iv].replace(",", ""))
    us = v / 1e3 if r[iu] in ("ns", "nsecond") else (v if r[iu] in ("us", "usecond") else v * 1e3)
    name = re.sub(r"\(.*", "", r[ik])[:72]
    agg[name][0] += 1
    agg[name][1] += us
tot = sum(v[1] for v in agg.values())
print("%-72s %7s %12s %7s" % ("kernel", "count", "total_us", "share"))
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-72s %7d %12.1f %6.1f%%" % (k, n, us, 100 * us / tot))
