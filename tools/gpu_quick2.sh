#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for cfg in "3 100000" "2 1000000"; do
set -- $cfg
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/hull_launches_d$1.csv python - $1 $2 <<'PY' > /dev/null 2>&1
import sys, numpy as np, hvb200
d, n = int(sys.argv[1]), int(sys.argv[2])
xs = np.random.default_rng(0).random((n, d))
cv = hvb200.ConvexHull(xs)
PY
python - gpurun_out/hull_launches_d$1.csv <<'PY'
import csv, collections, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    k = r[ki].split("(")[0][-40:]; v = float(r[vi].replace(",", "")); 
    if r[ui] == "ns": v /= 1e3
    elif r[ui] == "ms": v *= 1e3
    a = agg.setdefault(k, [0, 0.0, 0.0, []]); a[0] += 1; a[1] += v; a[2] = max(a[2], v); a[3].append(round(v,1))
for k, a in agg.items(): print("%-42s n=%4d total %10.1f us  max %9.1f us" % (k, a[0], a[1], a[2]), a[3][:12] if "wrap" in k else "")
PY
done
