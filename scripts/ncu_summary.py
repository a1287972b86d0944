"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, avg, share."""
import csv, collections, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
tot = collections.defaultdict(float); cnt = collections.Counter()
for row in csv.DictReader(lines):
    name = re.sub(r'\(.*', '', row['Kernel Name'])
    try:
        v = float(row['Metric Value'].replace(',', ''))
    except ValueError:
        continue
    unit = row['Metric Unit']
    v = v / 1e3 if unit == 'ns' else (v * 1e3 if unit == 'ms' else v)
    tot[name] += v; cnt[name] += 1
T = sum(tot.values())
print("launches %d, total %.0f us" % (sum(cnt.values()), T))
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print("%-36s n=%6d  total %10.0f us  avg %8.1f us  share %5.1f%%" % (k[:36], cnt[k], v, v / cnt[k], 100 * v / T))
