"""Sum an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

    python tools/launch_shares.py gpurun_out/launches.csv > profiles/rN_launch_shares.csv
"""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}
acc = OrderedDict()
for r in rows[1:]:
    name = re.sub(r'\(.*$', '', r[ik]).replace('emg::', '').replace('void ', '').strip()
    us = float(r[iv].replace(',', '')) * scale.get(r[iu], 1.0)
    n, t = acc.get(name, (0, 0.0))
    acc[name] = (n + 1, t + us)
total = sum(t for _, t in acc.values())
print("kernel,launches,total_us,share_pct,avg_us")
for name, (n, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print(f"\"{name}\",{n},{t:.1f},{100 * t / total:.2f},{t / n:.2f}")
