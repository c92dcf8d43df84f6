#!/usr/bin/env python3
"""Round-2 golden vectors, produced by the compiled reference (oracle/_ref, built from /root/reference by
oracle/Makefile) -- run in the build container, where the reference is present:

    python tests/golden/make_golden_r02.py        -> tests/golden/ref_vectors_r02.json

Kept apart from make_golden.py so that the round-1 fixtures stay byte-identical.  Contents:

  ntt12            NTT (ml_kem.c:287) on 12-bit inputs with coefficients in [q, 4096): the reference's butterfly leaves
                   the difference unreduced (:317-318).  Both builds of the reference (-O2 and the makefile's -g) agree.
  intt12_undefined InverseNTT (ml_kem.c:336) on such inputs: `Q - (t - f[j+len])` wraps in a 24-bit field and the product
                   with zeta overflows a signed int (:366-368).  The two builds DISAGREE; one input with both outputs is
                   recorded as evidence that there is no reference behaviour to reproduce.
  decaps_random    Decaps_internal / PKE_Decrypt on decapsulation keys made of random bytes (s^, t^ coefficients >= q, D4).
  addsub, vector_multiply
                   PolyAddition / PolySubtraction (ml_kem.c:580, :599) incl. 12-bit operands, VectorMultiply (:618), k = 2, 3, 4.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import REF_G_SO, REF_SO, Reference, build, sizes  # noqa: E402


def main():
    build()
    r, rg = Reference(REF_SO), Reference(REF_G_SO)
    rng = np.random.default_rng(20261018_02)
    v = {"generator": "tests/golden/make_golden_r02.py", "reference": "/root/reference ml_kem.c + sha3.c, gcc -O2 and gcc -g"}
    polys = []
    for t in range(12):
        f = np.zeros(256, np.uint16)
        if t < 4:      # the all-"difference" chains start at coefficients 0 and 1
            f[:2] = rng.integers(3329, 4096, 2)
        elif t < 6:
            f[:2] = rng.integers(3329, 4096, 2)
            f[2:] = rng.integers(0, 3, 254)
        elif t == 6:
            f[:] = 4095
        else:
            f = rng.integers(0, 4096, 256, dtype=np.uint16)
        a, b = r.ntt(f), rg.ntt(f)
        assert (a == b).all(), "the two builds of the reference must agree on NTT"
        polys.append({"f": f.tobytes().hex(), "ntt_f": a.tobytes().hex()})
    assert any((np.frombuffer(bytes.fromhex(p["ntt_f"]), np.uint16) >= 3329).any() for p in polys)
    v["ntt12"] = polys
    for _ in range(1000):
        f = rng.integers(0, 4096, 256, dtype=np.uint16)
        a, b = r.intt(f), rg.intt(f)
        if (a != b).any():
            v["intt12_undefined"] = {"f": f.tobytes().hex(), "gcc_O2": a.tobytes().hex(), "gcc_g": b.tobytes().hex()}
            break
    else:
        raise SystemExit("no disagreement found -- the toolchain changed?")
    dec = []
    for ps in (512, 768, 1024):
        sz = sizes(ps)
        for t in range(3):
            dk = rng.integers(0, 256, sz["dk"], dtype=np.uint8)
            if t == 1:
                dk[: sz["dk_pke"]] = 0xFF
            if t == 2:
                dk[sz["dk_pke"] : sz["dk_pke"] + 384 * sz["k"]] = 0xFF
            c = rng.integers(0, 256, sz["c"], dtype=np.uint8)
            K = r.decaps_internal(ps, dk.tobytes(), c.tobytes())
            assert K == rg.decaps_internal(ps, dk.tobytes(), c.tobytes())
            mp = r.pke_decrypt(ps, dk[: sz["dk_pke"]].tobytes(), c.tobytes())
            dec.append({"set": ps, "dk": dk.tobytes().hex(), "c": c.tobytes().hex(), "K": K.hex(), "m": mp.hex()})
    v["decaps_random"] = dec
    # PolyAddition / PolySubtraction (ml_kem.c:580, :599) on 12-bit operands, VectorMultiply (:618) for k = 2, 3, 4
    ring = []
    for t in range(4):
        u = rng.integers(0, 4096 if t >= 2 else 3329, 256, dtype=np.uint16)
        w = rng.integers(0, 4096 if t >= 2 else 3329, 256, dtype=np.uint16)
        a, b = r.poly_add(u, w), r.poly_sub(u, w)
        assert (a == rg.poly_add(u, w)).all() and (b == rg.poly_sub(u, w)).all()
        ring.append({"u": u.tobytes().hex(), "v": w.tobytes().hex(), "add": a.tobytes().hex(), "sub": b.tobytes().hex()})
    v["addsub"] = ring
    vm = []
    for k in (2, 3, 4):
        for hi in (3329, 4096):
            u = rng.integers(0, hi, (k, 256), dtype=np.uint16)
            w = rng.integers(0, 3329, (k, 256), dtype=np.uint16)
            o = r.vector_multiply(u, w, k)
            assert (o == rg.vector_multiply(u, w, k)).all()
            vm.append({"k": k, "u": u.tobytes().hex(), "v": w.tobytes().hex(), "w": o.tobytes().hex()})
    v["vector_multiply"] = vm
    out = os.path.join(ROOT, "tests", "golden", "ref_vectors_r02.json")
    json.dump(v, open(out, "w"), indent=0, sort_keys=True)
    print("wrote", out, os.path.getsize(out), "bytes", hashlib.sha256(open(out, "rb").read()).hexdigest()[:16])


if __name__ == "__main__":
    main()
