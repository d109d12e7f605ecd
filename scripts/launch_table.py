"""Per-kernel table of the LAST frame of an ncu launch list (gpu__time_duration.sum CSV of scripts/profile_static.py).
usage: launch_table.py <csv> <launches_per_frame>"""
import csv
import re
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
data = [r for r in rows[h + 1:] if len(r) > mv]
n = int(sys.argv[2])
last = data[-n:]
agg = OrderedDict()
tot = 0.0
for r in last:
    name = re.sub(r"\(.*", "", r[kn])
    name = re.sub(r"void |<unnamed>::|bevf::|\(anonymous namespace\)::", "", name)
    us = float(r[mv].replace(",", "")) / 1000.0
    tot += us
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-70s x%-3d %8.1f us  %5.1f %%" % (k[:70], c, us, 100 * us / tot))
print("total %.1f us over %d launches" % (tot, len(last)))
