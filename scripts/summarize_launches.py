"""CPU: summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, re, sys
rows = []
with open(sys.argv[1]) as f:
  lines = [l for l in f if not l.startswith('==')]
for r in csv.DictReader(lines):
  if r.get('Metric Name') != 'gpu__time_duration.sum':
    continue
  v = float(r['Metric Value'].replace(',', ''))
  unit = r.get('Metric Unit', 'ns')
  v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0}.get(unit, 1e-6)
  name = re.sub(r'\(.*', '', r['Kernel Name'])
  name = re.sub(r'^(void )?(dd::)?', '', name)
  rows.append((name, v))
tot = sum(v for _, v in rows)
by = collections.defaultdict(lambda: [0.0, 0])
for n, v in rows:
  by[n][0] += v; by[n][1] += 1
print(f'# {len(rows)} launches, {tot:.2f} ms summed (cold-cache, serialised under ncu: compare SHARES)')
for n, (v, c) in sorted(by.items(), key=lambda kv: -kv[1][0]):
  print(f'{v:8.3f} ms {100*v/tot:5.1f}%  x{c:5d}  avg {1000*v/c:8.1f} us  {n}')
