#!/usr/bin/env python3
"""Per-kernel times of keyed ML-KEM-768 Encaps + Decaps (expanded key table) on device-resident inputs.
MLKEM_B200_LIB selects an A/B build of the library."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import crystals_kyber_b200 as ck
from crystals_kyber_b200 import lib as L

if os.environ.get("MLKEM_B200_LIB"):
    L.load(os.path.abspath(os.environ["MLKEM_B200_LIB"]))
n = 1 << int(os.environ.get("LOG2N", "20"))
nk = 1 << int(os.environ.get("LOG2K", "16"))
kem = ck.MLKEM()
g = torch.Generator(device="cuda").manual_seed(1)
d, z = (torch.randint(0, 256, (nk, 32), dtype=torch.uint8, device="cuda", generator=g) for _ in range(2))
m = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
table = kem.keys_load(768, seeds=(d, z), expand=True)
kem.set_streams(int(os.environ.get("STREAMS", "1")))
for _ in range(2):
    c, K = kem.encaps_keyed(table, None, m)
    Kd = kem.decaps_keyed(table, None, c)
torch.cuda.synchronize()
assert bool((Kd == K).all())
kem.profile(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    c, K = kem.encaps_keyed(table, None, m)
    Kd = kem.decaps_keyed(table, None, c)
e1.record()
torch.cuda.synchronize()
rep = kem.profile_report()
kem.profile(False)
print(json.dumps({"lib": os.environ.get("MLKEM_B200_LIB", "product"), "n": n, "keys": nk, "pair_ms": e0.elapsed_time(e1) / 3,
                  "kernels_ms": {k[:44]: round(v["ms"] / 3, 3) for k, v in rep.items()}}))
