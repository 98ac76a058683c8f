"""Summarise an ncu launch list (gpu__time_duration.sum) of tools/time_lu.py by kernel name (tools, not product)."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5 and r[0].isdigit()]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4].split("(")[0]
    agg[name][0] += 1
    agg[name][1] += float(r[-1].replace(",", ""))
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:28s} {v[0]:6d} launches {v[1] / 1e6:9.3f} ms  {100 * v[1] / tot:5.1f} %")
print(f"{'total':28s} {sum(v[0] for v in agg.values()):6d} launches {tot / 1e6:9.3f} ms")
