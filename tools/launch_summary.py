"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals for ONE proof (the 2nd)."""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = [(x["Kernel Name"], float(x["Metric Value"].replace(",", ""))) for x in csv.DictReader(lines) if x.get("Metric Name") == "gpu__time_duration.sum"]
starts = [i for i, (k, _) in enumerate(rows) if "k_set_globals" in k]
a = starts[1] if len(starts) > 1 else starts[0]
b = starts[2] if len(starts) > 2 else len(rows)
agg = collections.OrderedDict(); tot = 0.0
for k, v in rows[a:b]:
    n = k.split("(")[0].replace("void ", "").replace("b200::", "")
    agg.setdefault(n, [0, 0.0]); agg[n][0] += 1; agg[n][1] += v; tot += v
print("one proof: %d launches, %.3f ms (serialised, cold-cache ncu timing: compare shares)" % (b - a, tot / 1e6))
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-44s n=%3d %9.3f ms %5.1f%%" % (k[:44], n, v / 1e6, 100 * v / tot))
