#!/usr/bin/env python3
"""What a caller of the reference gets: the latency of KEM_KeyGen / KEM_Encaps / KEM_Decaps through include/ml_kem.h (a batch of
one per call) and the throughput of small host-memory batches -- bench.py's `caller_view` on its own, for A/B runs:

    python tools/time_latency.py                               # small batches hash with one sponge per warp (default)
    MLKEM_B200_WARP_HASH_MAX=0 python tools/time_latency.py    # one sponge per thread everywhere (the form before)
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import crystals_kyber_b200 as ck

kem = ck.MLKEM()
out = bench.caller_view(kem)
print(json.dumps({"MLKEM_B200_WARP_HASH_MAX": os.environ.get("MLKEM_B200_WARP_HASH_MAX", "default (1024)"),
                  "drop_in_api_latency_us": {k: round(v["median"], 1) for k, v in out["drop_in_api_latency_us"].items()},
                  "batch_of_one": {k: v for k, v in out["batch_of_one"].items() if k != "note"},
                  "small_batch_pairs_per_s": {k: round(v["pairs_per_s"]) for k, v in out["small_batch_host_memory"].items()}}))
