#!/usr/bin/env python3
"""Per-kernel times of KeyGen_internal (BASELINE configs[2]) on device-resident seeds, all parameter sets."""
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
d, z = (torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g) for _ in range(2))
for ps in (512, 768, 1024):
    kem.set_streams(0)
    for _ in range(2):
        kem.keygen(ps, d, z)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        kem.keygen(ps, d, z)
    e1.record()
    torch.cuda.synchronize()
    overlapped = e0.elapsed_time(e1) / 3
    kem.set_streams(1)
    kem.keygen(ps, d, z)
    torch.cuda.synchronize()
    kem.profile(True)
    for _ in range(3):
        kem.keygen(ps, d, z)
    torch.cuda.synchronize()
    rep = kem.profile_report()
    kem.profile(False)
    print(json.dumps({"set": ps, "n": n, "keygen_ms": overlapped, "keys_per_s": n / overlapped * 1e3,
                      "serial_kernels_ms": {k[:40]: round(v["ms"] / 3, 3) for k, v in rep.items()}}))
