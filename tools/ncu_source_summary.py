#!/usr/bin/env python
"""Summarise the per-instruction stall sampling of one kernel from an ncu report.

    ncu -i report.ncu-rep --page source --csv > src.csv ; python tools/ncu_source_summary.py src.csv [--bars 0x38410]

Splits the SASS into warp roles at the markers a warp-specialised tcgen05 kernel has (TMA loads / UTCHMMA / LDTM),
then prints for each role: warp-instructions, samples, samples spent in mbarrier wait loops (per barrier offset) and
the remaining samples by stall reason. Written for csrc/enc_head.cu; the role split is a heuristic."""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path)))
    hdr, rr = rows[1], rows[2:]
    stall = [(i, c[6:]) for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
    num = lambda r, c: int(r[c]) if r[c].isdigit() else 0
    total = sum(num(r, 2) for r in rr)
    mma = [i for i, r in enumerate(rr) if "UTCHMMA" in r[1]]
    ldtm = [i for i, r in enumerate(rr) if "LDTM" in r[1]]
    tma = [i for i, r in enumerate(rr) if "UTMALDG" in r[1]]
    # role boundaries: front = after the last one-time TMA load up to the code that leads into the first MMA
    b0 = (tma[-1] + 1) if tma else 0
    # walk back from the first MMA to the previous unconditional branch / EXIT as the role boundary
    b1 = mma[0]
    while b1 > b0 and not re.match(r"\s*(BRA|EXIT|@!?P\d\s+EXIT)", rr[b1][1]):
        b1 -= 1
    b2 = ldtm[0]
    while b2 > mma[-1] and not re.match(r"\s*(BRA|EXIT)", rr[b2][1]):
        b2 -= 1
    regions = [("setup", 0, b0), ("front", b0, b1), ("mma", b1, b2), ("epilogue", b2, len(rr))]
    print("total samples %d, %d SASS rows; role boundaries %s" % (total, len(rr), [(n, a, b) for n, a, b in regions]))
    for name, a, b in regions:
        inst = sum(num(r, 5) for r in rr[a:b])
        samp = sum(num(r, 2) for r in rr[a:b])
        waits = collections.Counter()
        other = collections.Counter()
        for i in range(a, b):
            r = rr[i]
            m = re.search(r"TRYWAIT.*\+0x([0-9a-f]+)\]", r[1])
            in_wait = bool(m) or ("BRA" in r[1] and i > 0 and "TRYWAIT" in rr[i - 1][1])
            if m:
                key = m.group(1)
            elif in_wait:
                key = re.search(r"\+0x([0-9a-f]+)\]", rr[i - 1][1]).group(1)
            if in_wait:
                waits[key] += num(r, 2)
            else:
                for c, nm in stall:
                    other[nm] += num(r, c)
        print("%-9s warp-instr %10d  samples %6d (%.1f%%)  in mbarrier waits %6d %s" %
              (name, inst, samp, 100.0 * samp / max(total, 1), sum(waits.values()), dict(waits.most_common(5))))
        print("          other stalls: %s" % other.most_common(8))


if __name__ == "__main__":
    main()
