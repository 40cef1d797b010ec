"""Static SASS statistics of a kernel in libsfmloss.so: instruction count and opcode histogram of every loop
(backward branch target .. branch), so instruction budgets can be checked here before spending GPU time.
usage: python tools/sass_loop.py <substring of mangled kernel name> [lib]"""
import collections
import re
import subprocess
import sys

lib = sys.argv[2] if len(sys.argv) > 2 else 'sfm_learner_chainer_b200/libsfmloss.so'
pat = sys.argv[1]
out = subprocess.run(['cuobjdump', '-sass', lib], stdout=subprocess.PIPE, text=True).stdout
funcs = {}
name = None
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = m.group(1)
        funcs[name] = []
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', line)
    if m and name:
        funcs[name].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in funcs.items():
    if pat not in name:
        continue
    print('==', name, len(ins), 'instructions')
    addr_index = {a: k for k, (a, _) in enumerate(ins)}
    loops = []
    for k, (a, txt) in enumerate(ins):
        if 'BRA' not in txt:
            continue
        m = re.search(r'0x([0-9a-f]+)\s*$', txt)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr_index:
                loops.append((addr_index[tgt], k))
    for lo, hi in loops:
        body = ins[lo:hi + 1]
        c = collections.Counter()
        for _, txt in body:
            t = txt.split()
            op = t[1] if t[0].startswith('@') else t[0]
            base = op.split('.')[0]
            if base == 'MUFU':
                base = '.'.join(op.split('.')[:2])
            c[base] += 1
        n = len(body)
        fma = sum(c[k] for k in ('FFMA', 'FMUL', 'FADD', 'FFMA2', 'FMUL2', 'FADD2', 'IMAD', 'HFMA2'))
        print('  loop %#x..%#x: %d instr (fma-pipe %d, other %d)' % (ins[lo][0], ins[hi][0], n, fma, n - fma))
        print('    ' + ', '.join('%s %d' % kv for kv in c.most_common(40)))
