#!/usr/bin/env python3
"""Per-kernel times of ML-KEM-768 Encaps on device-resident inputs, no result checks.

Used with the experiment build (make exp; MLKEM_B200_LIB=build/libmlkem_b200_exp.so MLKEM_B200_EXPERIMENT=<bits>)
to measure what a phase of the fused kernel costs by leaving it out (the results are then wrong on purpose)."""
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
d, z, m = (torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g) for _ in range(3))
ek, dk = kem.keygen(768, d, z)
kem.set_streams(int(os.environ.get("STREAMS", "1")))
for _ in range(2):
    kem.encaps(768, ek, m)
torch.cuda.synchronize()
kem.profile(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    kem.encaps(768, ek, m)
e1.record()
torch.cuda.synchronize()
rep = kem.profile_report()
kem.profile(False)
print(json.dumps({"experiment": os.environ.get("MLKEM_B200_EXPERIMENT", "0"), "n": n, "encaps_ms": e0.elapsed_time(e1) / 3,
                  "kernels_ms": {k[:40]: round(v["ms"] / 3, 3) for k, v in rep.items()}}))
