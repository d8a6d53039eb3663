"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count / average / total (us)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, agg = None, collections.defaultdict(list)
for r in rows:
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get("Metric Name") == "gpu__time_duration.sum":
            agg[d["Kernel Name"][:70]].append(float(d["Metric Value"].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%-72s n=%4d  avg=%10.1f us  total=%10.1f us  %5.1f%%" % (k, len(v), sum(v) / len(v) / 1e3, sum(v) / 1e3, 100 * sum(v) / tot))
