#!/usr/bin/env python3
"""Runs every stand-alone batched primitive a few times on 2^18 polynomials (device memory): the command
profiled by ncu for the NTT / codec kernels (profiles/ncu_primitives_*.csv)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import crystals_kyber_b200 as ck

kem = ck.MLKEM()
n = 1 << 18
g = torch.Generator(device="cuda").manual_seed(1)
f = torch.randint(0, 3329, (n, 256), generator=g, device="cuda", dtype=torch.int16).view(torch.uint16)
h = torch.randint(0, 3329, (n, 256), generator=g, device="cuda", dtype=torch.int16).view(torch.uint16)
seeds = torch.randint(0, 256, (n, 34), generator=g, device="cuda", dtype=torch.uint8)
for _ in range(3):
    fh = kem.ntt(f)
    kem.intt(fh)
    kem.multiply_ntts(fh, h)
    for d in (4, 10, 12):
        b = kem.compress_encode(f, d)
        kem.decode_decompress(b, d)
    kem.sample_ntt(seeds)
torch.cuda.synchronize()
print("ok")
