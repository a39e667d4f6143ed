#!/usr/bin/env python
"""Static schedule of the basic blocks of one kernel, from `cuobjdump -sass` of an object / shared library.

    python tools/sass_sched.py ast_sac_b200/csrc/kernels_fast.o 'k_envILi0ELi1ELi0ELi0E' [--dump N]

Decodes the control word of every sm_100a instruction (bits 105..125 of the 128-bit encoding: stall count 4 bits,
yield 1, write barrier 3, read barrier 3, wait mask 6) and prints, per basic block of at least --min instructions, the
instruction count, the FP64 instructions (DFMA / DMUL / DADD / DSETP) and the sum of the stall counts -- the cycles one
warp alone needs to issue the block when no scoreboard wait binds.  --dump N prints the N-th listed block.
"""
import argparse
import re
import subprocess


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, name = None, None
    res = {}
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            name = m.group(1)
            cur = res.setdefault(name, [])
            continue
        if cur is not None:
            cur.append(ln)
    return res


def parse(lines):
    """-> list of (addr, text, stall, yield, wbar, rbar, wait) and the set of label addresses"""
    ins = []
    i = 0
    while i < len(lines):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?)\s*;?\s*/\* 0x([0-9a-f]{16}) \*/", lines[i])
        if m and i + 1 < len(lines):
            m2 = re.match(r"\s*/\* 0x([0-9a-f]{16}) \*/", lines[i + 1])
            if m2:
                hi = int(m2.group(1), 16)
                ins.append((int(m.group(1), 16), m.group(2).rstrip(" ;"), (hi >> 41) & 0xf, (hi >> 45) & 1,
                            (hi >> 46) & 7, (hi >> 49) & 7, (hi >> 52) & 0x3f))
                i += 2
                continue
        i += 1
    return ins


def blocks(ins):
    targets = set()
    for a, t, *_ in ins:
        m = re.search(r"\b(?:BRA|BSSY|BSSY\.\w+|CALL\.REL\.NOINC|BRA\.\w+)\b.*?(0x[0-9a-f]+)\s*$", t)
        if m and ("BRA" in t or "BSSY" in t or "CALL" in t):
            targets.add(int(m.group(1), 16))
    out, cur = [], []
    for rec in ins:
        a, t = rec[0], rec[1]
        if a in targets and cur:
            out.append(cur)
            cur = []
        cur.append(rec)
        op = t.split()[1] if t.startswith("@") and len(t.split()) > 1 else t.split()[0]
        if re.match(r"(BRA|EXIT|RET|CALL|BSYNC|BREAK|WARPSYNC|JMP|BRX)", op):
            out.append(cur)
            cur = []
    if cur:
        out.append(cur)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("path")
    ap.add_argument("kernel")
    ap.add_argument("--min", type=int, default=60)
    ap.add_argument("--dump", type=int, default=-1)
    a = ap.parse_args()
    ks = kernels(a.path)
    names = [k for k in ks if a.kernel in k]
    assert names, f"no kernel matching {a.kernel}; have {list(ks)[:8]} ..."
    for name in names:
        ins = parse(ks[name])
        print(f"{name}: {len(ins)} instructions")
        n = 0
        for b in blocks(ins):
            if len(b) < a.min:
                continue
            fp64 = sum(1 for r in b if re.search(r"\b(DFMA|DMUL|DADD|DSETP)", r[1]))
            print(f"  block {n}: 0x{b[0][0]:05x}..0x{b[-1][0]:05x}  {len(b)} instr, {fp64} FP64, stall sum "
                  f"{sum(r[2] for r in b)}")
            if n == a.dump:
                for (ad, t, st, y, wb, rb, wt) in b:
                    w = "".join(str(i) for i in range(6) if wt >> i & 1)
                    print(f"    {ad:05x} s{st}{' Y' if y else '  '} w{wb if wb != 7 else '-'} r{rb if rb != 7 else '-'} "
                          f"wait[{w:6s}] {t}")
            n += 1


if __name__ == "__main__":
    main()
