#!/usr/bin/env python
"""Aggregate an ncu SASS source-page CSV by CUDA source line.

    cuobjdump -xelf all libshipenv.so ; nvdisasm -g -c *.cubin > dis.txt
    ncu -i prof.ncu-rep --page source --csv --kernel-name regex:<k> --launch-count 1 > src.csv
    python tools/ncu_by_line.py dis.txt src.csv <mangled kernel substring> [file substring]

Instructions of inlined library code (CUDA math headers) are attributed to the last line of the
given file that preceded them in the instruction stream.
"""
import collections
import csv
import re
import sys


def line_map(dis_path, kernel, fname):
    """offset -> line number of `fname`."""
    m, cur, active = {}, None, False
    for ln in open(dis_path, errors="replace"):
        if ln.startswith("//---") and ".text." in ln:
            active = kernel in ln
            cur = None
            continue
        if not active:
            continue
        mm = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if mm:
            if fname in mm.group(1):
                cur = int(mm.group(2))
            continue
        mm = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if mm:
            m[int(mm.group(1), 16)] = cur
    return m


def main():
    dis, src, kernel = sys.argv[1:4]
    fname = sys.argv[4] if len(sys.argv) > 4 else "shipenv_kernels.cuh"
    lm = line_map(dis, kernel, fname)
    rows = list(csv.reader(open(src)))
    hdr = rows[1]
    ia, ismp, iaddr, ith = (hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Address"),
                            hdr.index("Thread Instructions Executed"))
    data = [r for r in rows[2:] if len(r) == len(hdr) and r[iaddr].startswith("0x")]
    base = min(int(r[iaddr], 16) for r in data)
    inst, smp, thr = collections.Counter(), collections.Counter(), collections.Counter()
    for r in data:
        line = lm.get(int(r[iaddr], 16) - base)
        inst[line] += float(r[ia] or 0)
        smp[line] += float(r[ismp] or 0)
        thr[line] += float(r[ith] or 0)
    ti, ts = sum(inst.values()), sum(smp.values())
    text = open([a for a in sys.argv[5:6]][0]).read().splitlines() if len(sys.argv) > 5 else None
    print(f"total warp instructions {ti:.0f}, stall samples {ts:.0f}")
    print(f"{'line':>6} {'inst%':>7} {'smp%':>7} {'lanes':>6}  source")
    for line, c in inst.most_common(60):
        srcline = text[line - 1].strip()[:90] if text and line else ""
        lanes = thr[line] / c if c else 0
        print(f"{str(line):>6} {100 * c / ti:7.2f} {100 * smp[line] / ts:7.2f} {lanes:6.1f}  {srcline}")


if __name__ == "__main__":
    main()
