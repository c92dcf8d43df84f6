#!/usr/bin/env python3
"""A/B timing of the stand-alone transforms and of the Decaps kernels on device-resident inputs
(MLKEM_B200_LIB selects the library build).  No result checks: parity is the job of tests/."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import crystals_kyber_b200 as ck
from crystals_kyber_b200 import lib as L

if os.environ.get("MLKEM_B200_LIB"):
    L.load(os.path.abspath(os.environ["MLKEM_B200_LIB"]))
n = 1 << int(os.environ.get("LOG2N", "20"))
kem = ck.MLKEM()
g = torch.Generator(device="cuda").manual_seed(1)
f = torch.randint(0, 3329, (n, 256), dtype=torch.int32, device="cuda", generator=g).to(torch.uint16)

def timed(fn, reps=5):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

out = {"lib": os.environ.get("MLKEM_B200_LIB", "default"), "n": n}
out["ntt_Gpolys"] = n / timed(lambda: kem.ntt(f)) / 1e6
out["intt_Gpolys"] = n / timed(lambda: kem.intt(f)) / 1e6
d, z, m = (torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g) for _ in range(3))
ek, dk = kem.keygen(768, d, z)
c, K = kem.encaps(768, ek, m)
kem.set_streams(1)
kem.decaps(768, dk, c)
torch.cuda.synchronize()
kem.profile(True)
for _ in range(3):
    kem.decaps(768, dk, c)
torch.cuda.synchronize()
rep = kem.profile_report()
kem.profile(False)
out["decaps_kernels_ms"] = {k[:34]: round(v["ms"] / 3, 3) for k, v in rep.items()}
print(json.dumps(out))
