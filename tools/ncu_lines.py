"""Developer script: attribute the per-SASS-instruction metrics of an ncu report to CUDA source lines.
usage: ncu_lines.py report.ncu-rep object.o kernel_substring [top]"""
import re, csv, subprocess, collections, sys, os, tempfile
rep, obj, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
dis = subprocess.run(['nvdisasm', '-g', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split('\n')
seq = []; cur = None; on = False
for l in dis:
    if l.startswith('.text.'):
        on = kname in l; continue
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: seq.append((cur, m.group(2).strip()))
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split('\n')))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) > 5]
iss = hdr.index('Warp Stall Sampling (All Samples)'); iex = hdr.index('Instructions Executed')
print('sass', len(seq), 'profiled', len(data))
agg = collections.defaultdict(lambda: [0, 0, 0])
for (loc, sass), r in zip(seq, data):
    a = agg[loc]; a[0] += int(r[iex]); a[1] += int(r[iss]); a[2] += 1
tot = sum(a[0] for a in agg.values()) or 1; tots = sum(a[1] for a in agg.values()) or 1
srcdir = os.path.dirname(os.path.abspath(obj)); src = {}
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    f, l = loc if loc else ('?', 0)
    if f not in src:
        try: src[f] = open(os.path.join(srcdir, f)).read().split('\n')
        except OSError: src[f] = None
    text = src[f][l - 1].strip()[:110] if src[f] else ''
    print('%5.1f%% stall %5.1f%% ex  n=%4d %s:%d  %s' % (100 * a[1] / tots, 100 * a[0] / tot, a[2], f, l, text))
