"""Share of every kernel in an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_shares.py file.csv [top]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 16
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[idx["Metric Value"]].replace(",", ""))
    u = r[idx["Metric Unit"]]
    v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
    k = r[idx["Kernel Name"]][:96]
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print("total us", round(tot), "launches", sum(v[0] for v in agg.values()))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"  {100 * v[1] / tot:5.1f}% {v[0]:5d} {v[1] / v[0]:8.1f}us {k}")
