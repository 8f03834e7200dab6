"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel share of the summed kernel time.
usage: summarize_launches.py launches.csv steps_in_run > summary.txt"""
import csv, re, sys
path, steps = sys.argv[1], float(sys.argv[2])
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(r['Metric Value'].replace(',', ''))
    unit = r['Metric Unit']
    ms = v / 1e6 if unit in ('nsecond', 'ns') else v / 1e3 if unit in ('usecond', 'us') else v
    name = re.sub(r'\(anonymous namespace\)::|unnamed>::|mvd::|void ', '', r['Kernel Name'])
    name = re.sub(r'\(.*$', '', name)[:80]
    rows.append((name, ms))
tot = sum(ms for _, ms in rows)
agg = {}
for n, ms in rows:
    a = agg.setdefault(n, [0.0, 0])
    a[0] += ms; a[1] += 1
print(f'ncu launch list summary: {path}, {len(rows)} launches over {steps:g} training steps')
print(f'(per-launch times are cold-cache and serialised under ncu: compare SHARES). total {tot / steps:.2f} ms of kernel time per step')
print('  ms/step  share launches/step  kernel')
for n, (ms, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f'{ms / steps:9.3f} {100 * ms / tot:5.1f}% {c / steps:13.1f}  {n}')
