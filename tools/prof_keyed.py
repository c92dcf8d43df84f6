#!/usr/bin/env python3
"""The ncu target for the keyed path: a table of 2^12 keys with expanded matrices, then ROUNDS x (keyed Encaps + keyed
Decaps) of ML-KEM-768 on 2^LOG2N device-resident items, chunks serialised on one stream."""
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
nk = 1 << 12
d, z = (torch.randint(0, 256, (nk, 32), dtype=torch.uint8, device="cuda", generator=g) for _ in range(2))
m = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
table = kem.keys_load(768, seeds=(d, z), expand=True)
l1 = kem.launch_count()
for r in range(rounds):
    c, K = kem.encaps_keyed(table, None, m)
    Kd = kem.decaps_keyed(table, None, c)
    if r == 0:
        l2 = kem.launch_count()
torch.cuda.synchronize()
assert bool((Kd == K).all())
print(f"items {n}  launches per round {l2 - l1}")
