"""Aggregate an ncu per-launch CSV (gpu__time_duration.sum) by kernel name."""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    k = row['Kernel Name'][:70]; v = float(row['Metric Value']); u = row['Metric Unit']
    v = v / 1000 if u == 'ns' else (v * 1000 if u == 'ms' else v)
    agg.setdefault(k, []).append(v)
for k, v in agg.items():
    if len(sys.argv) > 2 and sys.argv[2] not in k: continue
    print(f"{k:70s} n={len(v):3d} avg={sum(v)/len(v):9.1f}us min={min(v):.1f} max={max(v):.1f}")
