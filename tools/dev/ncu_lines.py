"""Per-source-line instruction and stall-sample totals from `ncu -i rep --page source --csv --print-source cuda,sass` (development aid)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = collections.OrderedDict()
fname = ''
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        fname = r[1].split('/')[-1]
    if len(r) > 8 and r[0] not in ('', 'Line No'):
        try:
            key = (fname, int(r[0]))
        except ValueError:
            continue
        def f(x):
            try:
                return float(x)
            except ValueError:
                return 0.0
        inst = f(r[7]); samp = f(r[4])
        a = agg.setdefault(key, [r[1].strip(), 0.0, 0.0])
        a[1] += inst; a[2] += samp
tot_i = sum(v[1] for v in agg.values()); tot_s = sum(v[2] for v in agg.values())
print(f"total inst {tot_i:.3e} samples {tot_s:.0f}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{k[0]}:{k[1]:5d} inst {100*v[1]/tot_i:5.1f}% stall {100*v[2]/tot_s:5.1f}%  {v[0][:110]}")
