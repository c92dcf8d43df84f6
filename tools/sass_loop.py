#!/usr/bin/env python3
"""Opcode histogram of the innermost loop bodies of one kernel (cuobjdump -sass): every backward branch closes a loop
[target, branch].  Used to pin "one Keccak round = 122 LOP3 + 58 SHF" on the two-round loop body of a hash kernel.

    python tools/sass_loop.py crystals-kyber_b200/libmlkem_b200.so 'k_decaps_G<.*3, 2, 2, 10, 4>'
"""
import collections, re, subprocess, sys


def main():
    so, pat = sys.argv[1], re.compile(sys.argv[2])
    txt = subprocess.run(f"cuobjdump -sass {so} | c++filt", shell=True, capture_output=True, text=True).stdout
    cur, insts = None, collections.defaultdict(list)
    for line in txt.splitlines():
        m = re.search(r"Function : (.*)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);", line)
        if m and cur and pat.search(cur):
            insts[cur].append((int(m.group(1), 16), m.group(3), m.group(4)))
    for k, lst in insts.items():
        print(k[:140])
        for addr, op, rest in lst:
            if op.startswith("BRA"):
                t = re.search(r"0x([0-9a-f]+)", rest)
                if t and int(t.group(1), 16) < addr:
                    lo = int(t.group(1), 16)
                    body = [o for a, o, _ in lst if lo <= a <= addr]
                    c = collections.Counter(o.split(".")[0] if not o.startswith("IMAD") else o for o in body)
                    print(f"  loop 0x{lo:x}..0x{addr:x}: {len(body)} instructions  " + "  ".join(f"{o}:{n}" for o, n in c.most_common(12)))


if __name__ == "__main__":
    main()
