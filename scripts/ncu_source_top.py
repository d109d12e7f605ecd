"""Top source lines by warp-stall samples from `ncu -i rep --page source --csv --kernel-name <k>` output (stdin or file)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = None
for i, r in enumerate(rows):
    if "Source" in r and any("Sampling" in c for c in r):
        hdr = i
        break
if hdr is None:
    print("no source table found; columns:", rows[0][:20] if rows else None)
    sys.exit(0)
h = rows[hdr]
src = h.index("Source")
samp = next(i for i, c in enumerate(h) if c.startswith("# Samples") or "Warp Stall Sampling (All" in c)
inst = next((i for i, c in enumerate(h) if c.startswith("Instructions Executed")), None)
data = []
for r in rows[hdr + 1:]:
    if len(r) <= samp:
        continue
    try:
        data.append((float(r[samp] or 0), float(r[inst] or 0) if inst is not None else 0, r[src].strip()))
    except ValueError:
        continue
tot = sum(d[0] for d in data) or 1.0
toti = sum(d[1] for d in data) or 1.0
for s, ins, line in sorted(data, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print("%5.1f %% samples  %5.1f %% inst   %s" % (100 * s / tot, 100 * ins / toti, line[:140]))
