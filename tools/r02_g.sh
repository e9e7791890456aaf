#!/bin/bash
O=gpurun_out; mkdir -p $O
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/g_launches_c5ldos.csv python tools/bench_dos.py C5 2048 1 1 200 ldos > $O/g_ncu.log 2>&1
python tools/launch_breakdown.py $O/g_launches_c5ldos.csv 2>/dev/null || python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/g_launches_c5ldos.csv")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]; ki = h.index("Kernel Name"); vi = h.index("Metric Value")
agg = collections.Counter(); cnt = collections.Counter()
for r in rows[hi + 1:]:
    try: v = float(r[vi].replace(",", ""))
    except Exception: continue
    nm = r[ki].split("(")[0][:60]; agg[nm] += v; cnt[nm] += 1
for k, v in agg.most_common(10): print(f"{k:60s} {cnt[k]:3d} {v/1e6:10.3f} ms")
PY
