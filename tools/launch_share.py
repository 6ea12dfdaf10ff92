"""Share of time per kernel from an ncu launch list (gpu__time_duration csv): python tools/launch_share.py file.csv"""
import csv, sys, collections, re
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki])
    t = float(r[vi].replace(",", ""))
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += t
tot = sum(a[1] for a in agg.values())
unit = rows[1][hdr.index("Metric Unit")]
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{100 * t / tot:5.1f} %  {n:4d} x {t / n:10.1f} {unit}  {name[:120]}")
print("total", tot, unit)
