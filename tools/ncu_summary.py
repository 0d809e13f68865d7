"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total time, share."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 5]
hdr, agg = None, collections.OrderedDict()
for r in rows:
    if r[0] == "ID":
        hdr = r
        continue
    if hdr is None:
        continue
    d = dict(zip(hdr, r))
    try:
        v = float(d["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    v = v / 1e3 if d["Metric Unit"] == "ns" else (v * 1e3 if d["Metric Unit"] == "ms" else v)
    name = d["Kernel Name"].split("(")[0][-70:]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches (cold-cache, serialised: compare shares)")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:10.1f} us {n:4d} launches {100 * t / tot:5.1f}%  {k}")
