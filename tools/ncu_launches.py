"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares."""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
agg = {}
for r in rows[1:]:
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    scale = {'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3, 'nsecond': 1e-3}.get(r[ui], 1e-3)
    agg.setdefault(r[ki], []).append(v * scale)
tot = sum(sum(v) for v in agg.values())
print(f"total {tot:.1f} us over {sum(len(v) for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{sum(v):9.1f} us  n={len(v):3d}  avg {sum(v)/len(v):8.1f} us  {100*sum(v)/tot:5.1f}%  {k[:90]}")
