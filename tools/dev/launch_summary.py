"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel (development aid)."""
import csv, collections, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    try:
        v = float(row['Metric Value'].replace(',', ''))
    except Exception:
        continue
    u = row['Metric Unit']
    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else (v * 1e6 if u == 's' else v))
    short = re.sub(r'\(.*', '', row['Kernel Name'])
    short = re.sub(r'^void ', '', short).replace('rolo::<unnamed>::', '').replace('<unnamed>::', '')[:60]
    agg[short][0] += 1; agg[short][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:60s} n={v[0]:5d} total={v[1]:11.1f} us avg={v[1]/v[0]:10.1f} us share={100*v[1]/tot:5.1f}%")
print(f"{'TOTAL':60s} {tot:.1f} us")
