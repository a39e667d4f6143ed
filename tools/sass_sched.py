#!/usr/bin/env python
"""Static schedule of a kernel's SASS: decodes the control word of every instruction (stall count, yield, write /
read barrier, wait mask) from `cuobjdump -sass` output so that dependent-issue chains can be read off the listing.

usage: tools/sass_sched.py file.sass [lo_hex hi_hex]
"""
import re
import sys


def parse(path):
    lines = open(path).read().split('\n')
    out = []
    i = 0
    while i < len(lines):
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/', lines[i])
        if m and i + 1 < len(lines):
            m2 = re.match(r'\s+/\* (0x[0-9a-f]{16}) \*/', lines[i + 1])
            if m2:
                hi = int(m2.group(1), 16)
                ctrl = hi >> 41
                stall = ctrl & 0xf
                yld = (ctrl >> 4) & 1
                wbar = (ctrl >> 5) & 7
                rbar = (ctrl >> 8) & 7
                wait = (ctrl >> 11) & 0x3f
                out.append((int(m.group(1), 16), m.group(2).strip(), stall, yld, wbar, rbar, wait))
                i += 2
                continue
        i += 1
    return out


if __name__ == "__main__":
    ins = parse(sys.argv[1])
    lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 60
    tot = 0
    for a, t, stall, yld, wbar, rbar, wait in ins:
        if lo <= a <= hi:
            tot += stall
            w = ''.join(str(b) for b in range(6) if wait >> b & 1)
            print(f"{a:05x} s{stall:<2d}{'Y' if yld else ' '} w{wbar if wbar != 7 else '-'} r{rbar if rbar != 7 else '-'} wait[{w:6s}] {t}")
    print("sum of stall counts in range:", tot)


def blocks(ins):
    """Basic blocks (split at branches / reconvergence points): (start, n_instr, n_fp64, sum of stall counts)."""
    out, cur = [], []
    for rec in ins:
        cur.append(rec)
        op = rec[1].split()[1] if rec[1].startswith('@') else rec[1].split()[0]
        if op.split('.')[0] in ('BRA', 'BSYNC', 'EXIT', 'RET', 'CALL', 'BREAK', 'WARPSYNC'):
            out.append(cur)
            cur = []
    if cur:
        out.append(cur)
    res = []
    for b in out:
        fp64 = sum(1 for r in b if re.search(r'\b(DFMA|DMUL|DADD|DSETP)\b', r[1]))
        res.append((b[0][0], len(b), fp64, sum(r[2] for r in b)))
    return res
