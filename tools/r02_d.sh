#!/bin/bash
O=gpurun_out; mkdir -p $O
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/d_launches_heev512.csv python tools/bench_heev.py 512 148 vec > $O/d_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/d_launches_heev512.csv")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]; ki = h.index("Kernel Name"); vi = h.index("Metric Value")
agg = collections.Counter(); cnt = collections.Counter()
for r in rows[hi + 1:]:
    try: v = float(r[vi].replace(",", ""))
    except Exception: continue
    nm = r[ki].split("(")[0][:60]; agg[nm] += v; cnt[nm] += 1
for k, v in agg.most_common(8): print(f"{k:60s} {cnt[k]:3d} {v/1e6:10.3f} ms")
PY
python tools/bench_vec.py C5 592
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee $O/d_pytest.log
