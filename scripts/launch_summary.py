"""Per-kernel summary of an ncu launch list (--metrics gpu__time_duration.sum --csv): launches, total and mean time, share.
usage: python scripts/launch_summary.py gpurun_out/X.csv [title]"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
tot, cnt = collections.defaultdict(float), collections.Counter()
for r in rows:
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
    v = float(r["Metric Value"].replace(",", ""))
    v = v / 1e3 if r["Metric Unit"] == "ns" else v
    tot[name] += v
    cnt[name] += 1
all_us = sum(tot.values())
print("%s: %d launches, %.0f us of kernel time (serialised, cold caches: compare shares)" % (
    sys.argv[2] if len(sys.argv) > 2 else sys.argv[1], len(rows), all_us))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print("%-28s n=%6d total %9.0f us avg %7.1f us share %5.1f%%" % (k, cnt[k], v, v / cnt[k], 100 * v / all_us))
