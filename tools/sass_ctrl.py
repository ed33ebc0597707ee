"""Decodes the scheduling control fields (stall, yield, write/read barrier, wait mask) of the SASS of one kernel between
two addresses.  usage: cuobjdump -sass lib.so | python tools/sass_ctrl.py <kernel-substring> <start-hex> <end-hex>"""
import re
import sys

want, lo, hi = sys.argv[1], int(sys.argv[2], 16), int(sys.argv[3], 16)
on = False
cur = None
for line in sys.stdin:
    if 'Function :' in line:
        on = want in line
        continue
    if not on:
        continue
    m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/', line)
    if m:
        cur = (int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16))
        continue
    m = re.match(r'\s+/\* (0x[0-9a-f]+) \*/', line)
    if m and cur:
        addr, text, w0 = cur
        w1 = int(m.group(1), 16)
        cur = None
        if not (lo <= addr <= hi):
            continue
        ctrl = w1 >> 41  # bits 105.. of the 128-bit word
        stall = ctrl & 0xf
        yld = (ctrl >> 4) & 1
        wbar = (ctrl >> 5) & 7
        rbar = (ctrl >> 8) & 7
        wait = (ctrl >> 11) & 0x3f
        print('%05x st%-2d %s W%s R%s wait[%s] %s' % (addr, stall, 'Y' if yld else ' ', '-' if wbar == 7 else wbar,
                                                     '-' if rbar == 7 else rbar,
                                                     ''.join(str(i) if wait >> i & 1 else '.' for i in range(6)), text[:90]))
