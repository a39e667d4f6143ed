#!/usr/bin/env python
"""Dynamic opcode mix of a kernel from `ncu -i X.ncu-rep --page source --csv --print-source sass` output.

usage: tools/ncu_opmix.py src.csv [section_index] [--lines]
Prints executed warp instructions per opcode, normalised by the execution count of the simulator loop's head
(the most frequent non-zero count), with the stall samples attributed to each opcode.
"""
import collections
import csv
import re
import sys


def sections(path):
    rows = list(csv.reader(open(path)))
    out, cur, hdr = [], None, None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = []
            out.append((r[1], cur))
        elif r and r[0] == 'Address':
            hdr = r
        elif cur is not None and hdr is not None and len(r) == len(hdr):
            cur.append(r)
    return hdr, out


def main():
    path = sys.argv[1]
    sec = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 0
    hdr, secs = sections(path)
    name, data = secs[sec]
    isrc, iex = hdr.index("Source"), hdr.index("Instructions Executed")
    ith = hdr.index("Thread Instructions Executed")
    istall = hdr.index("Warp Stall Sampling (All Samples)")
    ex = [int(r[iex]) for r in data]
    cnt = collections.Counter(e for e in ex if e > 0)
    base = max(cnt.items(), key=lambda kv: kv[0] * kv[1])[0]     # the count that carries the most instructions
    print(name[:100])
    print("instructions", len(data), "loop count", base, "executed", sum(ex), "per loop iteration", round(sum(ex) / base, 1),
          "avg threads", round(sum(int(r[ith]) for r in data) / max(1, sum(ex)), 2))
    c, st = collections.Counter(), collections.Counter()
    for r, e in zip(data, ex):
        t = re.sub(r'^@!?U?P\d+\s+', '', r[isrc].strip())
        op = t.split()[0].split('.')[0]
        c[op] += e
        st[op] += int(r[istall])
    tot, tots = sum(c.values()), max(1, sum(st.values()))
    for k, v in c.most_common(40):
        print(f"{k:10s} {v / base:7.1f} {100 * v / tot:5.1f}%  stall {100 * st[k] / tots:5.1f}%")
    if "--lines" in sys.argv:
        for r, e in zip(data, ex):
            print(f"{e / base:6.2f} {int(r[istall]):6d}  {r[isrc].strip()}")


if __name__ == "__main__":
    main()
