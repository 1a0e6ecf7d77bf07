"""List the hottest SASS instructions (warp-stall samples) of an ncu source page:
   ncu -i x.ncu-rep --page source --csv > src.csv ; python profiles/ncu_hot.py src.csv [min_pct]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = rows[1]
ci, si = hdr.index("Source"), hdr.index("# Samples")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[2:] if len(r) > si and r[si]]
tot = sum(float(r[si]) for r in body)
print("total samples", tot)
for n, r in enumerate(body):
    s = float(r[si])
    if s > tot * minpct / 100:
        top = sorted(((float(r[i] or 0), hdr[i]) for i in stalls), reverse=True)[:2]
        print("%5d %7.0f %5.1f%%  %-70s %s" % (n, s, 100 * s / tot, r[ci].strip()[:70], " ".join("%s=%d" % (h, v) for v, h in top)))
