"""aggregate an ncu source page (cuda,sass csv) by CUDA source line: samples and warp instructions"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; hdr = None; agg = []
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) == 2: continue
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr and r and r[0] != '':
        try:
            samples = int(r[hdr.index('# Samples')]); inst = int(r[hdr.index('Instructions Executed')])
        except Exception:
            continue
        agg.append((samples, inst, cur, r[0], r[1].strip()[:100]))
tot = sum(a[0] for a in agg); toti = sum(a[1] for a in agg)
print('total samples', tot, 'total warp-instr', toti)
for a in sorted(agg, reverse=True)[:top]:
    print('%6d (%4.1f%%) inst=%9d (%4.1f%%) %s:%s  %s' % (a[0], 100 * a[0] / tot, a[1], 100 * a[1] / toti, a[2], a[3], a[4]))
