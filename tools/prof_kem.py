#!/usr/bin/env python3
"""The ncu target for the KEM kernels: KeyGen, then ROUNDS x (Encaps + Decaps) of ML-KEM-768 on 2^LOG2N device-resident
items, chunks serialised on one stream.  Prints the number of kernel launches of the set-up and of one round, so that
`ncu --launch-skip <setup + round> --launch-count <round>` captures exactly the second round (warm caches, warm clocks)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import crystals_kyber_b200 as ck

n = 1 << int(os.environ.get("LOG2N", "16"))
rounds = int(os.environ.get("ROUNDS", "2"))
kem = ck.MLKEM()
kem.set_streams(1)
g = torch.Generator(device="cuda").manual_seed(1)
d, z, m = (torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g) for _ in range(3))
l0 = kem.launch_count()
ek, dk = kem.keygen(768, d, z)
l1 = kem.launch_count()
for r in range(rounds):
    c, K = kem.encaps(768, ek, m)
    c[3::10, 5] ^= 1
    Kd = kem.decaps(768, dk, c)
    if r == 0:
        l2 = kem.launch_count()
torch.cuda.synchronize()
ok = (Kd == K).all(dim=1)
assert bool(ok[0]) and not bool(ok[3])
print(f"items {n}  setup launches {l1 - l0}  launches per round {l2 - l1} (library kernels only; torch's indexing kernels come on top)")
