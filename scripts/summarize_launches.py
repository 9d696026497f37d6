"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel total time and share."""
import csv, sys, re, collections
path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = collections.OrderedDict()
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"oat::", "", name)
    name = re.sub(r"\(.*$", "", name)[:100]
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    d = tot.setdefault(name, [0.0, 0])
    d[0] += us; d[1] += 1
total = sum(v[0] for v in tot.values())
n = sum(v[1] for v in tot.values())
print("total_us %.1f launches %d" % (total, n))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print("%10.1f us %5.1f%%  n=%4d  avg %8.1f us  %s" % (v[0], 100 * v[0] / total, v[1], v[0] / v[1], k))
