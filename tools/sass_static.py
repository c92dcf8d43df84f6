#!/usr/bin/env python3
"""Static opcode histogram of one kernel of a cubin / .so (cuobjdump -sass), split by issue pipe.

    python tools/sass_static.py crystals-kyber_b200/libmlkem_b200.so 'k_sample_matvec.*3, 2, 2, 10, 4>, 1'
"""
import collections, re, subprocess, sys
ALU = ("LOP3", "SHF", "IADD3", "PRMT", "VIMNMX", "VIADDMNMX", "ISETP", "SEL", "LEA", "MOV", "SGXT", "BMSK", "PLOP3", "VIADD", "IADD", "LOP", "FLO", "POPC", "IABS")
FMA = ("IMAD",)
def main():
    so, pat = sys.argv[1], re.compile(sys.argv[2])
    txt = subprocess.run(f"cuobjdump -sass {so} | c++filt", shell=True, capture_output=True, text=True).stdout
    cur, hist = None, collections.defaultdict(collections.Counter)
    for line in txt.splitlines():
        m = re.search(r"Function : (.*)", line)
        if m: cur = m.group(1); continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur and pat.search(cur):
            op = m.group(2); b = op.split(".")[0]
            if b == "IMAD": b = "IMAD.HI" if ".HI" in op else "IMAD.WIDE" if ".WIDE" in op else op if any(x in op for x in (".MOV", ".SHL", ".IADD")) else "IMAD"
            hist[cur][b] += 1
    for k, c in hist.items():
        alu = sum(n for o, n in c.items() if o.split(".")[0] in ALU); fma = sum(n for o, n in c.items() if o.startswith("IMAD"))
        print(k[:140]); print(f"  total {sum(c.values())}  alu {alu}  fma {fma}")
        print("  " + "  ".join(f"{o}:{n}" for o, n in c.most_common(28)))
if __name__ == "__main__":
    main()
