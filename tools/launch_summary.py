"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
cols = rows[hdr]
ki, vi, ui = cols.index('Kernel Name'), cols.index('Metric Value'), cols.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(',', ''))
    v = v / 1000 if r[ui] == 'ns' else (v * 1000 if r[ui] == 'ms' else v)
    agg[r[ki][:64]][0] += 1
    agg[r[ki][:64]][1] += v
tot = sum(v[1] for v in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:66s} n={n:5d} total={t:10.1f}us avg={t / n:8.2f}us share={t / tot:.3f}")
