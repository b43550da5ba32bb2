import csv, sys, collections
shapes = [tuple(map(int, l.split())) for l in open(sys.argv[2])]
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith('==')]
t = []
for row in csv.DictReader(lines):
    if 'zgemm' in row['Kernel Name'] and row['Metric Name'] == 'gpu__time_duration.sum':
        v = float(row['Metric Value'].replace(',', '')); t.append(v * {'ns': 1, 'us': 1e3, 'ms': 1e6}[row['Metric Unit']])
shapes = shapes[-len(t):]          # the profiled step is the last one
assert len(shapes) == len(t), (len(shapes), len(t))
cls = collections.defaultdict(lambda: [0, 0.0, 0.0])
for (m, n, k, up, nz), ns in zip(shapes, t):
    fl = 8.0 * m * n * k * (0.5 if up and m == n else 1.0)
    if up and m != n: fl = 8.0 * k * (m * n - m * m / 2)      # row block of an upper update
    if k <= 64: kc = "K<=64"
    elif k <= 512: kc = "K<=512"
    else: kc = "K>512"
    mc = "m<=64" if m <= 64 else ("m<=512" if m <= 512 else "m>512")
    key = (kc, mc, "split" if nz > 1 else "")
    c = cls[key]; c[0] += 1; c[1] += ns; c[2] += fl
tot_t = sum(c[1] for c in cls.values()); tot_f = sum(c[2] for c in cls.values())
print(f"total {tot_t/1e6:.1f} ms {tot_f:.3e} flop -> {tot_f/tot_t/1e3:.1f} TF")
for key, c in sorted(cls.items(), key=lambda x: -x[1][1]):
    print(f"{str(key):34s} n={c[0]:5d} {c[1]/1e6:8.2f} ms {c[2]:.3e} flop {c[2]/c[1]/1e3:6.1f} TF  lost vs 35TF: {(c[1]-c[2]/35e3)/1e6:6.2f} ms")
