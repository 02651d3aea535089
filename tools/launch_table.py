"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total microseconds, share."""
import csv, sys, collections, re
path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
agg = collections.OrderedDict()
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = val / 1000.0 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1000.0)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total µs | share |\n|---|---:|---:|---:|")
for name, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.1f | %.1f %% |" % (re.sub(r"\s+", " ", name)[:100], c, us, 100 * us / tot))
print("\nTotal %.0f µs" % tot)
