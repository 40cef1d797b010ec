"""Summarise an `ncu --page source --csv` dump: executed warp-instructions by opcode and hot regions."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
ops = collections.Counter(); tot = 0; samples = collections.Counter()
body = rows[2:]
for r in body:
    try: n = int(r[iex])
    except Exception: continue
    s = r[isrc].strip()
    parts = s.split()
    op = parts[1] if parts and parts[0].startswith('@') and len(parts) > 1 else (parts[0] if parts else '?')
    op = op.split('.')[0]
    ops[op] += n; tot += n
    try: samples[op] += int(r[ismp])
    except Exception: pass
print('total warp-instructions', tot)
for op, n in ops.most_common(28):
    print('%-10s %12d %5.1f%%   stall-samples %d' % (op, n, 100.0 * n / tot, samples[op]))
