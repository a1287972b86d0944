#!/bin/bash
# Per-kernel time of one objective evaluation (m = 2600) for a few settings (ncu launch lists).
# usage: eval_breakdown.sh "ENV=.. ENV=.." "ENV=.." ...
for cfg in "$@"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  env $cfg timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ev_$tag.csv python scripts/one_eval.py > /dev/null 2>&1
  python - "$tag" "$cfg" <<'PY'
import csv, collections, re, sys
lines = [l for l in open("gpurun_out/ev_%s.csv" % sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
idx = [i for i, r in enumerate(rows) if r["Kernel Name"].startswith("k_build")]
tot = collections.defaultdict(float)
for r in rows[idx[-1]:]:
    name = re.sub(r"\(.*", "", r["Kernel Name"]); v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
    tot[name] += v / 1e3 if u == "ns" else v
print("%s: total %.0f us" % (sys.argv[2], sum(tot.values())), {k: round(v) for k, v in tot.items() if v > 100})
PY
done
