#!/usr/bin/env python3
"""bench_sweep.py -- BASELINE.json configs[4]: batch-size sweep 2^16 .. 2^24 of the full ML-KEM-768 KEM
(KeyGen -> Encaps -> Decaps, 10 % tampered) at 1/2/4/8 GPUs.

    python bench_sweep.py [--min-log2 16] [--max-log2 24] [--reps 3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench_sweep.py

The batch size is the TOTAL over all ranks (strong scaling at fixed batch, as the config asks for a sweep of
the batch); every rank takes the contiguous shard `shard_range(B, rank, N)` of the global index range, inputs
come from the global index, there is no collective on the data path.  One JSON line per batch size:
full-KEM round trips per second (device-resident, CUDA events, max over ranks) plus the three phases.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min-log2", type=int, default=16)
    ap.add_argument("--max-log2", type=int, default=24)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--set", type=int, default=768)
    ap.add_argument("--no-cpu-reference", action="store_true", help="skip the reference's own CPU figure (last line, rank 0)")
    args = ap.parse_args()

    import torch

    import crystals_kyber_b200 as ck
    from crystals_kyber_b200 import workload as wl

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    kem = ck.MLKEM()
    ps = args.set

    def tmax(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sync():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for lg in range(args.min_log2, args.max_log2 + 1, 2):
        B = 1 << lg
        b, e = ck.shard_range(B, rank, world)
        d, z, m = wl.derive_inputs(lambda msg, ln: kem.hash_batch(1, msg, ln), b, e, dev)
        times = {"keygen": [], "encaps": [], "decaps": []}
        for rep in range(args.reps + 1):  # first repetition is the warm-up
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            sync()
            ev[0].record()
            ek, dk = kem.keygen(ps, d, z)
            ev[1].record()
            c, K = kem.encaps(ps, ek, m)
            ev[2].record()
            sel = wl.tamper_inplace(c, b)  # (a handful of tiny torch kernels; included in the decaps interval)
            Kd = kem.decaps(ps, dk, c)
            ev[3].record()
            sync()
            if rep:
                for i, k in enumerate(times):
                    times[k].append(tmax(ev[i].elapsed_time(ev[i + 1])))
        same = (Kd == K).all(dim=1)
        ok = torch.ones(e - b, dtype=torch.bool, device=dev)
        ok[sel] = False
        assert bool(same[ok].all()) and not bool(same[~ok].any())
        best = {k: min(v) for k, v in times.items()}
        total = sum(best.values())
        if rank == 0:
            print(json.dumps({"config": f"ML-KEM-{ps} full KEM (KeyGen+Encaps+Decaps), 10% tampered", "n_gpus": world, "log2_batch": lg,
                              "batch_total": B, "items_per_gpu": e - b, "round_trips_per_s": B / (total * 1e-3),
                              "keygen_per_s": B / (best["keygen"] * 1e-3), "encaps_per_s": B / (best["encaps"] * 1e-3),
                              "decaps_per_s": B / (best["decaps"] * 1e-3), "ms": best}), flush=True)
        del d, z, m, ek, dk, c, K, Kd
        torch.cuda.empty_cache()
    if rank == 0 and not args.no_cpu_reference:
        print(json.dumps(cpu_reference(ps)), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def cpu_reference(ps, per_thread=8):
    """BASELINE configs[4] asks for the host-core CPU reference beside the sweep: full KEM round trips per second of the
    unmodified reference (oracle/_ref, built by oracle/Makefile) on all host threads, on a bounded sample -- the
    reference's time per item does not depend on the batch size."""
    import numpy as np

    from oracle.oracle import REF_SO, Reference, build

    build()
    if not os.path.exists(REF_SO):
        return {"cpu_reference": None, "why": "oracle/_ref did not travel"}
    ref = Reference(REF_SO)
    threads = os.cpu_count() or 1
    n = threads * per_thread
    rng = np.random.default_rng(20261018)
    d, z, m = (rng.integers(0, 256, (n, 32), dtype=np.uint8) for _ in range(3))
    tk, ek, dk = ref.time_keygen(ps, d, z, threads)
    tp, _, _ = ref.time_pairs(ps, ek, dk, m, threads)
    return {"cpu_reference": f"ML-KEM-{ps} full KEM (KeyGen+Encaps+Decaps), the unmodified reference at gcc -O2", "cores": threads,
            "sample": f"{n} round trips on {threads} threads", "round_trips_per_s": n / (tk + tp), "keygen_per_s": n / tk,
            "encaps_plus_decaps_pairs_per_s": n / tp}


if __name__ == "__main__":
    main()
