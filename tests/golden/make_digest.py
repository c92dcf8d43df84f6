#!/usr/bin/env python3
"""Oracle-derived digests of BASELINE configs[3] over a fixed global range (CPU only; minutes on a few cores):

    python tests/golden/make_digest.py [log2_items=20]   -> tests/golden/config4_digest.json

Items are global indices [0, 2^L) of the workload of crystals-kyber_b200/workload.py (index-derived seeds, tamper rule
i % 10 == 3).  For each of c = Encaps output, K = encapsulated key, Kd = Decaps(tampered c) the digest is

    sha256( sha256(block 0) || sha256(block 1) || ... ),   block = 2^14 consecutive items,

a checksum of checksums that any contiguous sharding whose boundaries are multiples of 2^14 can compute shard by shard.
bench.py recomputes it on the GPUs at every rank count (1, 2, 4, 8) and compares: byte-identity of the concatenated
outputs at every G, against the oracle rather than against another GPU run."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "crystals-kyber_b200"))
import workload as wl  # noqa: E402  (plain module import: the package itself needs the CUDA library)
from oracle.oracle import Oracle, build  # noqa: E402

BLOCK_LOG2 = 14


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    build()
    orc = Oracle()
    n, blk = 1 << L, 1 << BLOCK_LOG2
    parts = {"c": [], "K": [], "Kd": []}
    for lo in range(0, n, blk):
        d, z, m = wl.derive_inputs(lambda msg, ln: orc.hash_batch(1, msg, ln), lo, lo + blk)
        ek, dk = orc.keygen(768, d, z)
        c, K = orc.encaps(768, ek, m)
        ct = c.copy()
        wl.tamper_inplace(ct, lo)
        Kd = orc.decaps(768, dk, ct)
        for name, a in (("c", c), ("K", K), ("Kd", Kd)):
            parts[name].append(wl.block_hashes(a, blk))
        print(f"\r{lo + blk}/{n}", end="", file=sys.stderr)
    out = {"param_set": 768, "log2_items": L, "log2_block": BLOCK_LOG2, "generator": "tests/golden/make_digest.py (oracle/mlkem_oracle.c)",
           "digest": {k: wl.combine_block_hashes(v) for k, v in parts.items()}}
    path = os.path.join(ROOT, "tests", "golden", "config4_digest.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print("\nwrote", path, out["digest"])


if __name__ == "__main__":
    main()
