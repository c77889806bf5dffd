"""Developer script: print the kernels of one frame, in launch order, from an ncu launch list (csv of gpu__time_duration.sum)."""
import csv, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]; ki = h.index('Kernel Name'); mi = h.index('Metric Value'); gi = h.index('Grid Size'); bi = h.index('Block Size'); si = h.index('Stream')
data = rows[1:]
idx = [i for i, r in enumerate(data) if 'k_filter_count' in r[ki]]
which = int(sys.argv[2]) if len(sys.argv) > 2 else -2
st = idx[which]; en = idx[which + 1] if which + 1 < 0 or which + 1 < len(idx) else len(data)
tot = 0.0
for r in data[st:en]:
    name = r[ki].replace('osep::', '').split('(')[0]
    us = float(r[mi]) / 1000.0; tot += us
    print(name[:40].ljust(40), r[gi].ljust(14), r[bi].ljust(14), r[si].ljust(4), "%8.1f" % us)
print("sum %.1f us" % tot)
