"""CPU tests of the multi-GPU host logic: contiguous index shards, inputs derived from the
global index, no collective on the data path.  The N > 1 path runs as a world-size-2 gloo group; the
per-shard arithmetic is done by the oracle here (no GPU in this container)."""
import hashlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from crystals_kyber_b200 import shard_range
from crystals_kyber_b200 import workload as wl


def test_shard_range_partitions():
    for n in (0, 1, 7, 1 << 16, (1 << 22) + 3):
        for w in (1, 2, 4, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            assert max(e - b for b, e in parts) - min(e - b for b, e in parts) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_inputs_depend_on_global_index_only(oracle):
    G = lambda msg, ln: oracle.hash_batch(1, msg, ln)
    d, z, m = wl.derive_inputs(G, 0, 64)
    d2, z2, m2 = wl.derive_inputs(G, 32, 64)
    assert (d[32:] == d2).all() and (z[32:] == z2).all() and (m[32:] == m2).all()
    assert d.tobytes()[:32] == hashlib.sha3_512(b"mlkemkey" + (0).to_bytes(8, "little")).digest()[:32]
    assert m[5].tobytes() == hashlib.sha3_512(b"mlkemmsg" + (5).to_bytes(8, "little")).digest()[:32]


def test_tamper_rule():
    c = np.zeros((40, 1088), np.uint8)
    sel = wl.tamper_inplace(c, 100)  # global items 100..139
    assert list(sel + 100) == [103, 113, 123, 133]
    for i in sel:
        g = i + 100
        assert c[i, (g * 7919) % 1088] == 1 << (g % 8) and c[i].sum() == 1 << (g % 8)
    assert c[np.setdiff1d(np.arange(40), sel)].sum() == 0


def test_work_model_matches_survey():
    ops = wl.op_counts(3, 2, 10, 4)
    assert (ops["keygen"], ops["encaps"], ops["decaps"]) == (232224, 250624, 267200)
    assert wl.keccak_calls(3, 2, 10, 4) == {"keygen": 43, "encaps": 44, "decaps": 42, "encrypt": 34}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle import Oracle

    orc = Oracle()
    b, e = shard_range(n, rank, world)
    d, z, m = wl.derive_inputs(lambda msg, ln: orc.hash_batch(1, msg, ln), b, e)
    ek, dk = orc.keygen(768, d, z)
    c, K = orc.encaps(768, ek, m)
    wl.tamper_inplace(c, b)
    Kd = orc.decaps(768, dk, c)
    # the only cross-rank traffic: a checksum of the shard's outputs and the timing reduction
    digest = hashlib.sha256(ek.tobytes() + c.tobytes() + Kd.tobytes()).digest()
    gathered = [None] * world
    dist.all_gather_object(gathered, (b, e, digest, int((Kd == K).all(axis=1).sum())))
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # max-over-ranks timing, as in bench.py
    if rank == 0:
        q.put((gathered, float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_run_matches_single_rank(oracle):
    n = 60
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 2.0
    assert [g[:2] for g in gathered] == [(0, 30), (30, 60)]
    # single-rank run over the whole range: the shards' outputs are slices of it
    d, z, m = wl.derive_inputs(lambda msg, ln: oracle.hash_batch(1, msg, ln), 0, n)
    ek, dk = oracle.keygen(768, d, z)
    c, K = oracle.encaps(768, ek, m)
    wl.tamper_inplace(c, 0)
    Kd = oracle.decaps(768, dk, c)
    for b, e, digest, n_equal in gathered:
        assert hashlib.sha256(ek[b:e].tobytes() + c[b:e].tobytes() + Kd[b:e].tobytes()).digest() == digest
        assert n_equal == (e - b) - len([i for i in range(b, e) if i % 10 == 3])


def _digest_worker(rank, world, port, total, blk, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle import Oracle

    orc = Oracle()
    b, e = shard_range(total, rank, world)
    d, z, m = wl.derive_inputs(lambda msg, ln: orc.hash_batch(1, msg, ln), b, e)
    ek, dk = orc.keygen(768, d, z)
    c, K = orc.encaps(768, ek, m)
    ct = c.copy()
    wl.tamper_inplace(ct, b)
    Kd = orc.decaps(768, dk, ct)
    mine = {name: wl.block_hashes(a, blk) for name, a in (("c", c), ("K", K), ("Kd", Kd))}
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    if rank == 0:
        q.put({k: wl.combine_block_hashes([p[k] for p in parts]) for k in ("c", "K", "Kd")})
    dist.barrier()
    dist.destroy_process_group()


def test_checksum_of_checksums_is_independent_of_the_rank_count(oracle):
    """The cross-N digest of bench.py (crystals-kyber_b200/workload.py block_hashes / combine_block_hashes): two gloo ranks
    over global items [0, 256) in blocks of 64 give the digest of a single-process run -- and the committed fixture
    tests/golden/config4_digest.json is this very construction over 2^20 items in blocks of 2^14 (its first block is
    recomputed here)."""
    import json

    total, blk = 256, 64
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_digest_worker, args=(r, 2, port, total, blk, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    d, z, m = wl.derive_inputs(lambda msg, ln: oracle.hash_batch(1, msg, ln), 0, total)
    ek, dk = oracle.keygen(768, d, z)
    c, K = oracle.encaps(768, ek, m)
    ct = c.copy()
    wl.tamper_inplace(ct, 0)
    Kd = oracle.decaps(768, dk, ct)
    want = {name: wl.combine_block_hashes([wl.block_hashes(a, blk)]) for name, a in (("c", c), ("K", K), ("Kd", Kd))}
    assert got == want
    fixture = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config4_digest.json")))
    assert fixture["log2_items"] == 20 and fixture["log2_block"] == 14 and set(fixture["digest"]) == {"c", "K", "Kd"}
    with pytest.raises(AssertionError):
        wl.block_hashes(c[:100], blk)  # a shard must be a whole number of blocks
