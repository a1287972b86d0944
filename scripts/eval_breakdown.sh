#!/bin/bash
# Per-kernel time of one objective evaluation (m = 2600) for a few factorisation settings (ncu launch lists).
for cfg in "2 4" "1 4" "1 8" "2 8"; do set -- $cfg
  MOGP_TRAIL_ROWS=$1 MOGP_FACTOR_KB=$2 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ev_$1_$2.csv python scripts/one_eval.py > /dev/null 2>&1
  python - "$1" "$2" <<'PY'
import csv, collections, re, sys
rows_, kb = sys.argv[1], sys.argv[2]
lines = [l for l in open("gpurun_out/ev_%s_%s.csv" % (rows_, kb)) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
idx = [i for i, r in enumerate(rows) if r["Kernel Name"].startswith("k_build")]
tot = collections.defaultdict(float)
for r in rows[idx[-1]:]:
    name = re.sub(r"\(.*", "", r["Kernel Name"]); v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
    tot[name] += v / 1e3 if u == "ns" else v
print("row tiles per CTA %s, panels per group %s: total %.0f us" % (rows_, kb, sum(tot.values())), {k: round(v) for k, v in tot.items() if v > 100})
PY
done
