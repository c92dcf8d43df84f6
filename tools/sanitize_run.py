#!/usr/bin/env python3
"""Small end-to-end run of every kernel (all parameter sets, host and device memory), used under
compute-sanitizer (memcheck / racecheck):  compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import crystals_kyber_b200 as ck

rng = np.random.default_rng(0)
for fips in (False, True):
    kem = ck.MLKEM(fips203=fips, chunk_items=64)
    for ps in (512, 768, 1024):
        n = 161
        d, z, m = (rng.integers(0, 256, (n, 32), dtype=np.uint8) for _ in range(3))
        ek, dk = kem.keygen(ps, d, z)
        c, K = kem.encaps(ps, ek, m)
        c[3, 7] ^= 1
        Kd = kem.decaps(ps, dk, c)
        assert (Kd[:3] == K[:3]).all() and (Kd[3] != K[3]).any()
        td, tz, tm = (torch.from_numpy(x).cuda() for x in (d, z, m))
        ek2, dk2 = kem.keygen(ps, td, tz)
        c2, K2 = kem.encaps(ps, ek2, tm)
        Kd2 = kem.decaps(ps, dk2, c2)
        torch.cuda.synchronize()
        assert torch.equal(K2, Kd2) and (ek2.cpu().numpy() == ek).all()
        kem.pke_decrypt(ps, kem.pke_keygen(ps, d)[1], c)
        kem.check_dk(ps, dk)
kem = ck.MLKEM()
f = rng.integers(0, 3329, (77, 256), dtype=np.uint16)
kem.intt(kem.ntt(f))
kem.multiply_ntts(f, f)
kem.sample_ntt(rng.integers(0, 256, (77, 34), dtype=np.uint8), return_seeds=True)
ck.MLKEM(sample_group_limit=150).sample_ntt(rng.integers(0, 256, (200, 34), dtype=np.uint8))
for eta in (2, 3):
    kem.cbd(rng.integers(0, 256, (77, 64 * eta), dtype=np.uint8), eta)
    kem.prf_cbd(rng.integers(0, 256, (77, 32), dtype=np.uint8), rng.integers(0, 256, 77, dtype=np.uint8), eta)
for dd in (1, 4, 5, 10, 11, 12):
    b = kem.compress_encode(f, dd)
    kem.decode_decompress(b, dd)
    kem.byte_decode(kem.byte_encode(f & ((1 << dd) - 1), dd), dd)
    kem.decompress(kem.compress(f, dd), dd)
for which, ln in ((0, 1184), (1, 64), (2, 1120), (3, 1120)):
    kem.hash_batch(which, rng.integers(0, 256, (77, ln), dtype=np.uint8), ln)
kem.sha3_bits(rng.integers(0, 2, (33, 1085), dtype=np.uint8), [0, 1, 0, 0], 512, 256)
torch.cuda.synchronize()
print("sanitize_run ok")
