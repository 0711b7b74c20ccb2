"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if 'Kernel Name' in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get('Metric Name') == 'gpu__time_duration.sum':
            name = re.sub(r'^void ', '', re.sub(r'\(.*', '', d['Kernel Name']))
            v = float(d['Metric Value'].replace(',', ''))
            v = v / 1e3 if d['Metric Unit'] == 'ns' else (v * 1e3 if d['Metric Unit'] == 'ms' else v)
            agg[name][0] += 1
            agg[name][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{v[1] / 1e3:9.3f} ms {v[0]:5d}  {100 * v[1] / tot:5.1f}%  {v[1] / v[0]:9.1f} us/launch  {k[:100]}")
print(f"total {tot / 1e3:.3f} ms")
