"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, mean, share."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
start = next(i for i, r in enumerate(rows) if r[0] == "ID")
hdr = rows[start]
d = collections.defaultdict(list)
for r in rows[start + 1:]:
    rec = dict(zip(hdr, r))
    if rec.get("Metric Name") == "gpu__time_duration.sum":
        v = float(rec["Metric Value"].replace(",", ""))
        if rec.get("Metric Unit") == "us":
            v *= 1e3
        elif rec.get("Metric Unit") == "ms":
            v *= 1e6
        d[rec["Kernel Name"].split("(")[0]].append(v)
tot = sum(sum(v) for v in d.values())
print(f"total {tot / 1e3:.1f} us over {sum(len(v) for v in d.values())} launches")
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k[:60]:60s} n={len(v):5d} avg={sum(v) / len(v) / 1e3:9.2f} us  max={max(v) / 1e3:9.2f} us  share={sum(v) / tot:.3f}")
