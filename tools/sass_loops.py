"""Lists the loops (backward branches) of one kernel in a cuobjdump -sass dump with their instruction mix.
usage: cuobjdump -sass lib.so | python tools/sass_loops.py <kernel-name-substring> [min_len]"""
import re
import sys
from collections import Counter

want, min_len = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 100
ins, on = [], False
for line in sys.stdin:
    if 'Function :' in line:
        on = want in line
        if on:
            print(line.strip())
        continue
    if not on:
        continue
    m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);', line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr2i = {a: i for i, (a, _) in enumerate(ins)}


def op(t):
    return re.sub(r'^@!?U?P\d+\s+', '', t).split()[0].split('.')[0]


print('total', len(ins), dict(Counter(op(t) for _, t in ins).most_common(12)))
for a, t in ins:
    m = re.search(r'BRA\S*\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)', t)
    if m and int(m.group(1), 16) < a and int(m.group(1), 16) in addr2i:
        s = int(m.group(1), 16)
        body = ins[addr2i[s]:addr2i[a] + 1]
        if len(body) >= min_len:
            c = Counter(op(t2) for _, t2 in body)
            fp64 = c['DADD'] + c['DFMA'] + c['DMUL']
            print(f'loop {s:#x}..{a:#x} n={len(body)} fp64={fp64} other={len(body) - fp64}', dict(c.most_common(16)))
