"""Rough in-order issue model of one basic block of SASS (cuobjdump -sass text): register/predicate RAW dependencies,
fixed latencies per op class, one instruction issued per cycle at most, 2 cycles between instructions of the same pipe.
Usage: python tools/sass_sim.py file.sass first_line last_line   (1-based line numbers of the instruction list)"""
import re, sys
LAT = {'MUFU': 18, 'FFMA': 4.5, 'FMUL': 4.5, 'FADD': 4.5, 'FSEL': 4.5, 'FSETP': 4.5 + 4, 'FMNMX': 4.5, 'LOP3': 4.5, 'IADD3': 4.5, 'ISETP': 4.5 + 4,
       'PLOP3': 4.5 + 4, 'SEL': 4.5, 'MOV': 4.5, 'IMAD': 4.5, 'VIADD': 4.5, 'HFMA2': 4.5, 'LDS': 29, 'LDG': 300, 'SHFL': 25}
PIPE = {'FFMA': 'fma', 'FMUL': 'fma', 'FADD': 'fma', 'IMAD': 'fma', 'HFMA2': 'fma', 'MUFU': 'xu'}
def parse(line):
    m = re.match(r'\s*(?:/\*[0-9a-f]+\*/)?\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*(.*?);', line)
    if not m: return None
    pred, op, args = m.group(1), m.group(2), m.group(3)
    base = op.split('.')[0]
    toks = [t.strip() for t in args.split(',')] if args else []
    regs = lambda s: re.findall(r'\b(?:R\d+|P\d+|UR\d+|UP\d+)\b', s)
    dst, src = [], []
    nd = 1
    if base in ('FSETP', 'ISETP', 'PLOP3'): nd = 2
    if base == 'LOP3' and toks and toks[0].startswith('P'): nd = 2
    if base in ('BRA', 'BSSY', 'BSYNC', 'ST', 'STS', 'STG'): nd = 0
    for i, t in enumerate(toks):
        (dst if i < nd else src).extend(regs(t))
    if pred: src.extend(regs(pred))
    return base, dst, src, line.strip()
def main():
    lines = open(sys.argv[1]).read().splitlines()
    a, b = int(sys.argv[2]), int(sys.argv[3])
    ready = {}; pipe_free = {}; t = 0.0; rows = []
    for ln in lines[a - 1:b]:
        p = parse(ln)
        if not p: continue
        base, dst, src, txt = p
        if base in ('BSSY', 'BSYNC', 'NOP'): continue
        dep = max([ready.get(r, 0.0) for r in src if r not in ('PT', 'RZ')] + [0.0])
        pipe = PIPE.get(base, 'alu')
        issue = max(t + 1, dep, pipe_free.get(pipe, 0.0))
        stall = issue - (t + 1)
        t = issue
        pipe_free[pipe] = issue + 2
        lat = LAT.get(base, 4.5)
        for r in dst:
            if r not in ('PT', 'RZ'): ready[r] = issue + lat
        rows.append((issue, stall, txt[:90]))
    for issue, stall, txt in rows:
        print("%7.1f %5.1f  %s" % (issue, stall, txt) if '-v' in sys.argv else '', end='\n' if '-v' in sys.argv else '')
    print("instructions %d, modelled cycles %.0f" % (len(rows), t))
    big = sorted(rows, key=lambda r: -r[1])[:25]
    print("largest stalls:")
    for issue, stall, txt in sorted(big):
        print("%7.1f %5.1f  %s" % (issue, stall, txt))
if __name__ == '__main__':
    main()
