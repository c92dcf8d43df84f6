#!/usr/bin/env python3
"""Opcode histogram of a kernel from the ncu source page (SASS view), weighted by executed instructions.

    ncu -i prof.ncu-rep --page source --csv > src.csv ; python tools/sass_hist.py src.csv [kernel-substring]

Splits the kernel at the Keccak loop (the LOP3/SHF-dense region) so that the remaining alu-pipe work shows up.
"""
import csv, sys, collections, re
ALU = {"LOP3", "SHF", "IADD3", "IADD", "PRMT", "IMNMX", "VIMNMX", "VIMNMX3", "ISETP", "SEL", "LEA", "MOV", "SGXT", "BMSK", "PLOP3", "IABS", "FLO", "POPC", "LOP", "VIADD", "IADD32I"}
FMA = {"IMAD", "IMAD.HI", "IMAD.WIDE", "FFMA", "FMUL", "FADD"}
def main():
    path = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""
    rows = list(csv.reader(open(path)))
    kern = None; hdr = None; out = collections.defaultdict(lambda: collections.Counter())
    samples = collections.defaultdict(lambda: collections.Counter())
    for r in rows:
        if len(r) >= 2 and r[0] == "Kernel Name": kern = r[1]; hdr = None; continue
        if r and r[0] == "Address": hdr = r; continue
        if hdr is None or len(r) < len(hdr) - 2: continue
        if want and want not in kern: continue
        src = r[hdr.index("Source")].strip(); n = int(r[hdr.index("Instructions Executed")] or 0)
        s = int(r[hdr.index("# Samples")] or 0)
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
        if not m: continue
        op = m.group(2); base = op.split(".")[0]
        if base == "IMAD" and ".HI" in op: base = "IMAD.HI"
        elif base == "IMAD" and ".WIDE" in op: base = "IMAD.WIDE"
        elif base == "IMAD" and (".MOV" in op or ".SHL" in op or ".IADD" in op): base = op  # pseudo forms
        out[kern][base] += n; samples[kern][base] += s
    for k, c in out.items():
        tot = sum(c.values())
        print(k[:120], "total warp-inst", tot)
        for op, n in c.most_common(40):
            print(f"  {op:14s} {n:12d} {100*n/tot:6.2f}%  samples {samples[k][op]}")
if __name__ == "__main__":
    main()
