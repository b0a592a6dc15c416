"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[hi]; kn = hdr.index('Kernel Name'); mv = hdr.index('Metric Value'); mu = hdr.index('Metric Unit')
agg = collections.OrderedDict(); seq = []
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    name = r[kn].split('(')[0].replace('void ', '').replace('svn::', ''); v = float(r[mv].replace(',', '')); u = r[mu]
    v = v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v
    agg.setdefault(name, []).append(v); seq.append((name, v))
tot = sum(sum(v) for v in agg.values())
print(f"{'kernel':32s} {'n':>4s} {'total ms':>9s} {'share':>6s} {'mean us':>9s} {'min us':>8s} {'max us':>8s}")
for k, v in agg.items():
    print(f"{k:32s} {len(v):4d} {sum(v)/1e3:9.3f} {100*sum(v)/tot:5.1f}% {sum(v)/len(v):9.1f} {min(v):8.1f} {max(v):8.1f}")
print(f"total {tot/1e3:.3f} ms over {len(seq)} launches")
if len(sys.argv) > 2:
    print([round(v) for n, v in seq if sys.argv[2] in n])
