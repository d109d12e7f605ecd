"""Per-kernel table of the launches after the first N kernels of an ncu launch list.  usage: <csv> <skip_fraction 0..1>"""
import csv
import re
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
data = [r for r in rows[h + 1:] if len(r) > mv]
n = len(data)
last = data[int(n * float(sys.argv[2])):]
agg = OrderedDict()
tot = 0.0
for r in last:
    name = re.sub(r"\(.*", "", r[kn])
    name = re.sub(r"void |<unnamed>::|bevf::|\(anonymous namespace\)::|at::native::", "", name)
    us = float(r[mv].replace(",", "")) / 1000.0
    tot += us
    a = agg.setdefault(name[:90], [0, 0.0])
    a[0] += 1
    a[1] += us
for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%-92s x%-4d %9.1f us  %5.1f %%" % (k, c, us, 100 * us / tot))
print("total %.1f us over %d launches" % (tot, len(last)))
