#!/usr/bin/env python3
"""Runs the KEM of all parameter sets on the experiment build (make exp), whose three-block sampler traps on any
shared-memory store outside its slot -- the stand-in for compute-sanitizer, which is closed on the GPU pool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import crystals_kyber_b200 as ck
from crystals_kyber_b200 import lib as L

L.load(os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "build", "libmlkem_b200_exp.so")))
rng = np.random.default_rng(3)
for ps in (512, 768, 1024):
    for n in (1, 31, 50001):
        kem = ck.MLKEM()
        d, z, m = (torch.from_numpy(rng.integers(0, 256, (n, 32), dtype=np.uint8)).cuda() for _ in range(3))
        ek, dk = kem.keygen(ps, d, z)
        c, K = kem.encaps(ps, ek, m)
        Kd = kem.decaps(ps, dk, c)
        torch.cuda.synchronize()
        assert torch.equal(K, Kd), (ps, n)
print("check_bounds ok: no store of the three-block sampler left its slot")
