"""aggregate an ncu source page by CUDA source line: stall samples and warp instructions.
usage: python tools/ncu_lines.py <report.ncu-rep> <kernel regex> [top]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + kern], capture_output=True, text=True).stdout
cur = None; hdr = None; agg = {}; seen_kernel = 0
for r in csv.reader(io.StringIO(raw)):
    if len(r) >= 2 and r[0] == 'Kernel Name':
        seen_kernel += 1
        if seen_kernel > 1: break
        continue
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) >= 2 and r[0] == 'Function Name': continue
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr and r and r[0].isdigit():
        try:
            samples = int(r[hdr.index('# Samples')]); inst = int(r[hdr.index('Instructions Executed')])
        except Exception:
            continue
        key = (cur, r[0])
        a = agg.setdefault(key, [0, 0, r[1].strip()[:110]])
        a[0] += samples; a[1] += inst
tot = sum(a[0] for a in agg.values()) or 1; toti = sum(a[1] for a in agg.values()) or 1
print('total samples', tot, 'total warp-instr', toti)
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print('%6d (%4.1f%%) inst=%9d (%4.1f%%) %s:%s  %s' % (a[0], 100 * a[0] / tot, a[1], 100 * a[1] / toti, f, ln, a[2]))
