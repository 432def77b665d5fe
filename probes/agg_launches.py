"""Aggregate an ncu --csv launch list (gpu__time_duration.sum) by kernel name."""
import csv, sys, collections, re
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
    name = re.sub(r"<.*", "", r[ki].split("(")[0])
    agg[name][0] += 1; agg[name][1] += ms
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:50s} launches {v[0]:6d}  total {v[1]:9.3f} ms  {100*v[1]/tot:5.1f}%  avg {1e3*v[1]/v[0]:9.1f} us")
print("total", tot, "ms")
