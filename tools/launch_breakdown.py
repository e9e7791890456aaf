"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.  usage: launch_breakdown.py file.csv"""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4][:90]; t = float(r[-1].replace(",", ""))
    agg.setdefault(name, [0, 0.0]); agg[name][0] += 1; agg[name][1] += t
tot = sum(v[1] for v in agg.values())
for k, v in agg.items():
    print(f"{v[0]:4d} {v[1]/1e6:9.3f} ms {100*v[1]/tot:5.1f}%  {k}")
