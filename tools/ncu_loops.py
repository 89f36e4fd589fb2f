"""Share of executed instructions and stall samples per innermost SASS loop of one ncu capture.
    ncu -i X.ncu-rep --page source --csv --print-source sass > src.csv;  python tools/ncu_loops.py src.csv"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]; data = rows[2:]
ia, isrc, isamp, iex = h.index('Address'), h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
base = int(data[0][ia], 16)
ins = [(int(r[ia], 16) - base, r[isrc].strip(), int(r[isamp] or 0), int(r[iex] or 0)) for r in data]
tot_ex = sum(x[3] for x in ins); tot_s = sum(x[2] for x in ins)
warps = ins[0][3]
print('total executed', tot_ex, 'samples', tot_s, 'instructions', len(ins), 'warps', warps)
addr = {a: i for i, (a, _, _, _) in enumerate(ins)}
loops = []
for i, (a, t, _, _) in enumerate(ins):
    m = re.search(r'BRA.*0x([0-9a-f]+)', t)
    if m:
        tgt = int(m.group(1), 16) - base
        if tgt < a and tgt in addr: loops.append((addr[tgt], i))
covered = [False] * len(ins)
out = []
for s, e in sorted(loops, key=lambda x: x[1] - x[0]):
    if e - s > 700 or any(covered[s:e + 1]): continue
    for k in range(s, e + 1): covered[k] = True
    ex = sum(x[3] for x in ins[s:e + 1]); sm = sum(x[2] for x in ins[s:e + 1])
    if ex * 200 > tot_ex or sm * 200 > tot_s:
        body = [x[1] for x in ins[s:e + 1]]
        cnt = lambda p: sum(1 for t in body if re.search(p, t))
        out.append((ins[s][0], ins[e][0], e - s + 1, ex, sm, cnt(r'F(ADD|MUL|FMA)2'), cnt(r'\bF(ADD|MUL|FMA)\b'), cnt('MUFU'), cnt(r'\bD(FMA|MUL|ADD)\b')))
for a, b, n, ex, sm, f2, f1, mu, df in sorted(out):
    print(f'loop {a:#7x}-{b:#7x} n={n:4d} (F2 {f2:3d} F {f1:3d} MUFU {mu:2d} D {df:3d}) exec {100*ex/tot_ex:5.1f}% samples {100*sm/tot_s:5.1f}%  trips {ex/n/warps:7.1f}/warp')
rest_ex = sum(x[3] for i, x in enumerate(ins) if not covered[i]); rest_s = sum(x[2] for i, x in enumerate(ins) if not covered[i])
print(f'outside those loops: exec {100*rest_ex/tot_ex:.1f}% ({rest_ex/warps:.0f} per warp) samples {100*rest_s/tot_s:.1f}%')
