"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line.
usage: python scripts/ncu_lines.py dump.csv path/to/source.cuh [kernel_index] [top]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
src = open(sys.argv[2]).read().split('\n')
want = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
kernels, cur = [], None
for r in rows:
    if r and r[0] == 'Function Name':
        cur = {'name': r[1], 'hdr': None, 'rows': []}; kernels.append(cur)
    elif cur is not None and r and r[0] == 'Line No':
        cur['hdr'] = r
    elif cur is not None and cur['hdr'] is not None and len(r) > 10:
        cur['rows'].append(r)
k = kernels[want]
hdr = k['hdr']; iS = hdr.index('# Samples'); iI = hdr.index('Instructions Executed')
stall = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
agg = {}
for r in k['rows']:
    if r[0] != '' and r[2] == '-':
        ln = int(r[0]); a = agg.setdefault(ln, [0, 0, collections.Counter()])
        a[0] += int(r[iS] or 0); a[1] += int(r[iI] or 0)
        for i, h in stall: a[2][h] += int(r[i] or 0)
tot = sum(v[0] for v in agg.values()) or 1; toti = sum(v[1] for v in agg.values()) or 1
print(k['name'], "samples", tot, "warp-inst", toti)
for ln, (s_, i_, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    t3 = ", ".join(f"{h[6:]}:{v}" for h, v in st.most_common(3))
    print(f"{ln:4d} smp {100*s_/tot:5.1f}% inst {100*i_/toti:5.1f}% | {src[ln-1].strip()[:64]:64s} | {t3}")
