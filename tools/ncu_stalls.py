"""Per-SASS-instruction stall samples of an .ncu-rep (source page): totals per stall reason and the top
instructions, each with the producers of its source registers (to see what a long-scoreboard wait is on).
usage: python tools/ncu_stalls.py <report.ncu-rep> [top N]"""
import csv, re, subprocess, sys
rep = sys.argv[1]; top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'sass', '--csv'], stdout=subprocess.PIPE,
                     stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = rows[1]; data = rows[2:]
col = {h: i for i, h in enumerate(hdr)}
ins = [r[1].strip() for r in data]
reasons = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[2]) for r in data)
print('samples', tot, {h[6:]: sum(int(r[col[h]]) for r in data) for h in reasons if sum(int(r[col[h]]) for r in data) > tot / 200})
def lastdef(i, reg):
    n = int(reg[1:])
    for j in range(i - 1, max(i - 2500, 0), -1):
        m = re.match(r'(@!?U?P\d+\s+)?(\S+)\s+(?:PT, )?(R\d+)', ins[j])
        if not m: continue
        op = m.group(2); d = int(m.group(3)[1:])
        width = 4 if '.128' in op else (2 if '.64' in op else 1)
        if d <= n < d + width: return j, ins[j][:60]
    return None
order = sorted(range(len(data)), key=lambda i: -int(data[i][2]))[:top_n]
for i in sorted(order):
    r = data[i]
    best = max(reasons, key=lambda h: int(r[col[h]]))
    print('%5d %-58s smp %6s  %s %s' % (i, ins[i][:58], r[2], best[6:], r[col[best]]))
    for rg in re.findall(r'\bR\d+\b', ins[i])[1:]:
        d = lastdef(i, rg)
        if d and ('LD' in d[1] or 'MUFU' in d[1] or 'SHFL' in d[1]): print('          %s <- %d %s' % (rg, d[0], d[1]))
