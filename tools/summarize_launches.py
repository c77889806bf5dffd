import csv, re, sys, collections
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
tot = 0; agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum': continue
    name = re.sub(r'\(.*', '', row['Kernel Name']).replace('osep::', '').replace('void ', '')
    v = float(row['Metric Value'].replace(',', '')); unit = row['Metric Unit']
    v = v / 1000 if unit == 'ns' else (v * 1000 if unit == 'ms' else v)
    tot += v; a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
print("total us %.1f  (%d launches)" % (tot, sum(a[0] for a in agg.values())))
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-34s x%-3d %9.1f us  %5.1f%%" % (k[:34], c, v, 100 * v / tot))
