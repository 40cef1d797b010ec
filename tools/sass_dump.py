"""Print selected instructions (regex) of one kernel's SASS within an address range, with their position.
usage: python tools/sass_dump.py <kernel substring> <lo hex> <hi hex> [regex] [lib]"""
import re, subprocess, sys
pat, lo, hi = sys.argv[1], int(sys.argv[2], 16), int(sys.argv[3], 16)
rx = re.compile(sys.argv[4]) if len(sys.argv) > 4 else re.compile('.')
lib = sys.argv[5] if len(sys.argv) > 5 else 'sfm_learner_chainer_b200/libsfmloss.so'
out = subprocess.run(['cuobjdump', '-sass', lib], stdout=subprocess.PIPE, text=True).stdout
on = False; k = 0
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        on = pat in m.group(1); k = 0
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', line)
    if on and m:
        a = int(m.group(1), 16)
        if lo <= a <= hi:
            k += 1
            if rx.search(m.group(2)):
                print('%4d %#06x  %s' % (k, a, m.group(2).strip()[:110]))
