import csv, collections, re, sys
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith('==')]
tot = collections.defaultdict(float); cnt = collections.Counter()
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum': continue
    name = re.sub(r'\(.*', '', row['Kernel Name']).replace('mspde::<unnamed>::', '')
    v = float(row['Metric Value'].replace(',', '')); unit = row['Metric Unit']
    ns = v * {'ns': 1, 'us': 1e3, 'ms': 1e6}.get(unit, 1)
    tot[name] += ns; cnt[name] += 1
T = sum(tot.values())
print("total ms %.3f launches %d" % (T / 1e6, sum(cnt.values())))
for k, v in sorted(tot.items(), key=lambda x: -x[1])[:12]:
    print(f"{v/1e6:9.3f} ms  {100*v/T:5.1f}%  n={cnt[k]:5d}  avg {v/cnt[k]/1e3:8.1f} us  {k[:60]}")
