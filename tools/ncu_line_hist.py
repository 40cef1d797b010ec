"""Per-source-line executed warp-instructions from `ncu --page source --print-source cuda,sass --csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
cur = None; out = []; tot = 0
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]; continue
    if len(r) < 8 or r[0] in ('Line No', 'Function Name'): continue
    if r[2] != '-': continue                 # SASS row, already counted in its line row
    try: n = int(r[7]); smp = int(r[6])
    except Exception: continue
    out.append((n, smp, cur, r[0], r[1].strip()[:110])); tot += n
print('total', tot)
out.sort(key=lambda t: (t[2], int(t[3])))
for n, smp, f, ln, src in out:
    if n * 1000 >= tot * (1 if top > 0 else 0) and n > tot / 400:
        print('%5.1f%% %11d smp %6d  %s:%s  %s' % (100.0 * n / tot, n, smp, f, ln, src))
