"""Top stalled SASS instructions per launch from `ncu -i rep --page source --csv --print-source sass`.
usage: ncu_top.py file.csv [n_top] [launch_index]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
only = int(sys.argv[3]) if len(sys.argv) > 3 else None
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
for k in range(len(starts) - 1):
    if only is not None and k != only:
        continue
    sec = rows[starts[k]:starts[k + 1]]
    print("=" * 20, "launch", k, sec[0][1][:110])
    hdr = sec[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in sec[2:] if len(r) == len(hdr)]
    tot = sum(int(r[idx["# Samples"]] or 0) for r in data)
    print("total samples", tot, "instructions", len(data))
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]] or 0))[:n]:
        s = int(r[idx["# Samples"]])
        st = sorted(((int(r[idx[h]] or 0), h[6:]) for h in stalls), reverse=True)[:2]
        print(f"{s:6d} {100 * s / max(tot, 1):5.1f}%  exec={r[idx['Instructions Executed']]:>8s} {r[idx['Source']].strip()[:64]:64s} {st}")
    agg = {h[6:]: sum(int(r[idx[h]] or 0) for r in data) for h in stalls}
    print("stall totals:", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
