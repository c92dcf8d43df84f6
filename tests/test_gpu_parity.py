"""GPU parity tests (-m gpu): the CUDA library, called through its C ABI, against the oracle.

Every comparison is bit-exact (integer / byte work).  Inputs are seeded; sizes are chosen so that the
oracle finishes in seconds.  Nothing here reads /root/reference: the anchors are the oracle (pinned to
the reference by tests/test_oracle.py), the committed golden vectors, and -- when oracle/_ref travelled
to the box -- the compiled reference itself.
"""
import hashlib

import numpy as np
import pytest

from archive_drivers import DRIVERS

pytestmark = pytest.mark.gpu

SETS = (512, 768, 1024)


@pytest.fixture(scope="module")
def mlkem():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import crystals_kyber_b200 as ck

    return ck.MLKEM()


def h2a(hexstr, dtype):
    return np.frombuffer(bytes.fromhex(hexstr), dtype=dtype)


def rng_polys(rng, n, hi=3329):
    return rng.integers(0, hi, (n, 256), dtype=np.uint16)


# ------------------------------------------------------------------ the reference's own drivers
@pytest.mark.parametrize("name", sorted(DRIVERS))
def test_archive_stdout(mlkem, archive_stdout, name):
    """Each Test_Archive driver, re-stated in tests/archive_drivers.py, run on the CUDA library."""
    out = DRIVERS[name](mlkem).encode()
    assert hashlib.sha256(out).hexdigest() == archive_stdout[name]["sha256"]


# ------------------------------------------------------------------ golden vectors produced by the reference
def test_golden_codec(mlkem, ref_vectors):
    for d in range(1, 13):
        x = np.arange(4096, dtype=np.uint16)
        assert hashlib.sha256(mlkem.compress(x, d).tobytes()).hexdigest() == ref_vectors["compress_sha256"][str(d)]
        y = np.arange(1 << d, dtype=np.uint16)
        assert hashlib.sha256(mlkem.decompress(y, d).tobytes()).hexdigest() == ref_vectors["decompress_sha256"][str(d)]
    for rec in ref_vectors["encode"]:
        d = rec["d"]
        F, B = h2a(rec["F"], np.uint16), h2a(rec["B"], np.uint8)
        assert (mlkem.byte_decode(B, d)[0] == F).all()
        if not rec.get("decode_only"):
            assert (mlkem.byte_encode(F, d)[0] == B).all()


def test_golden_samplers_hashes_ring(mlkem, ref_vectors):
    seeds = np.stack([h2a(r["seed"], np.uint8) for r in ref_vectors["sample_ntt"]])
    a, after = mlkem.sample_ntt(seeds, return_seeds=True)
    assert (a == np.stack([h2a(r["a"], np.uint16) for r in ref_vectors["sample_ntt"]])).all()
    assert (after == np.stack([h2a(r["seed_after"], np.uint8) for r in ref_vectors["sample_ntt"]])).all()
    for rec in ref_vectors["cbd"]:
        assert (mlkem.cbd(h2a(rec["B"], np.uint8), rec["eta"])[0] == h2a(rec["f"], np.uint16)).all()
    for rec in ref_vectors["hash"]:
        if rec["fn"] == "PRF":
            continue
        data = h2a(rec["in"], np.uint8)
        if len(data) % 8 or len(data) == 0:
            continue  # the batched hash entry point takes multiples of 8 bytes (all KEM-path inputs)
        which = {"H": 0, "G": 1, "J": 2}[rec["fn"]]
        assert mlkem.hash_batch(which, data, len(data))[0].tobytes().hex() == rec["out"]
    for rec in ref_vectors["ring"]:
        f, g = h2a(rec["f"], np.uint16), h2a(rec["g"], np.uint16)
        if rec["ntt_f"]:
            assert (mlkem.ntt(f)[0] == h2a(rec["ntt_f"], np.uint16)).all()
            assert (mlkem.intt(f)[0] == h2a(rec["intt_f"], np.uint16)).all()
        assert (mlkem.multiply_ntts(f, g)[0] == h2a(rec["mul"], np.uint16)).all()


def test_golden_kem(mlkem, ref_vectors):
    sha = lambda a: hashlib.sha256(np.asarray(a).tobytes()).hexdigest()
    for rec in ref_vectors["kem"]:
        ps = rec["set"]
        d, z, m = (h2a(rec[k], np.uint8) for k in ("d", "z", "m"))
        ek, dk = mlkem.keygen(ps, d, z)
        assert sha(ek) == rec["ek_sha256"] and sha(dk) == rec["dk_sha256"]
        c, K = mlkem.encaps(ps, ek, m)
        assert sha(c) == rec["c_sha256"] and K.tobytes().hex() == rec["K"]
        assert mlkem.decaps(ps, dk, c).tobytes().hex() == rec["K"]
        bad = c.copy()
        bad[0, 5] ^= 1
        assert mlkem.decaps(ps, dk, bad).tobytes().hex() == rec["K_rej"]
        cp = mlkem.pke_encrypt(ps, ek, m, h2a(rec["pke_r"], np.uint8))
        assert sha(cp) == rec["pke_c_sha256"]
        assert (mlkem.pke_decrypt(ps, dk, cp, dk_stride=dk.shape[1]) == m).all()
    rec = ref_vectors["ek_all_ff_768"]  # D4: values >= q in the key pass through ByteDecode12
    c, K = mlkem.encaps(768, np.full(1184, 0xFF, np.uint8), h2a(rec["m"], np.uint8))
    assert sha(c) == rec["c_sha256"] and K.tobytes().hex() == rec["K"]


# ------------------------------------------------------------------ primitives vs oracle, seeded random inputs
@pytest.mark.parametrize("n", [1, 7, 8, 9, 255, 4096, 40000])
def test_ring_vs_oracle(mlkem, oracle, n):
    rng = np.random.default_rng(n)
    f, g = rng_polys(rng, n), rng_polys(rng, n)
    fh = mlkem.ntt(f)
    assert (fh == oracle.ntt(f)).all()
    assert (mlkem.intt(f) == oracle.intt(f)).all()
    assert (mlkem.intt(fh) == f).all()  # NTT_test08.c round trip
    g12 = rng_polys(rng, n, 4096)       # operands as ByteDecode12 hands them over (D4)
    assert (mlkem.multiply_ntts(g12, f) == oracle.multiply_ntts(g12, f)).all()
    assert (mlkem.multiply_ntts(f, g) == oracle.multiply_ntts(f, g)).all()


def test_ntt_extreme_inputs(mlkem, oracle):
    f = np.stack([np.zeros(256, np.uint16), np.full(256, 3328, np.uint16), np.arange(256, dtype=np.uint16) * 13,
                  np.eye(1, 256, 255, dtype=np.uint16)[0] * 3328])
    assert (mlkem.ntt(f) == oracle.ntt(f)).all()
    assert (mlkem.intt(f) == oracle.intt(f)).all()


@pytest.mark.parametrize("n", [1, 31, 33, 127, 129, 5000])
def test_samplers_vs_oracle(mlkem, oracle, n):
    rng = np.random.default_rng(100 + n)
    seeds = rng.integers(0, 256, (n, 34), dtype=np.uint8)
    assert (mlkem.sample_ntt(seeds) == oracle.sample_ntt(seeds)).all()
    for eta in (2, 3):
        data = rng.integers(0, 256, (n, 64 * eta), dtype=np.uint8)
        assert (mlkem.cbd(data, eta) == oracle.cbd(data, eta)).all()
        s32 = rng.integers(0, 256, (n, 32), dtype=np.uint8)
        nonces = rng.integers(0, 256, n, dtype=np.uint8)
        assert (mlkem.prf_cbd(s32, nonces, eta) == oracle.prf_cbd(s32, nonces, eta)).all()


def test_sample_ntt_restart_path(oracle):
    """ml_kem.c:221-242: give up after the group limit, bump B[32], B[33], start over.  Unreachable with the
    reference's limit (P < 1e-90), so both sides run with a lowered limit (test hook on both)."""
    import crystals_kyber_b200 as ck

    rng = np.random.default_rng(5)
    seeds = rng.integers(0, 256, (3000, 34), dtype=np.uint8)
    for limit in (150, 158, 170):
        gpu = ck.MLKEM(sample_group_limit=limit)
        a, after = gpu.sample_ntt(seeds, return_seeds=True)
        oracle.set_sample_group_limit(limit)
        try:
            ea, eafter = oracle.sample_ntt_with_seeds(seeds)
        finally:
            oracle.set_sample_group_limit(0)
        assert (eafter != seeds).any(), "the lowered limit must trigger restarts for the test to mean anything"
        assert (a == ea).all() and (after == eafter).all()


@pytest.mark.parametrize("d", [1, 4, 5, 10, 11, 12])
def test_codec_vs_oracle(mlkem, oracle, d):
    rng = np.random.default_rng(d)
    n = 3001
    F = rng.integers(0, 1 << d, (n, 256), dtype=np.uint16)
    B = mlkem.byte_encode(F, d)
    assert (B == oracle.byte_encode(F, d)).all()
    assert (mlkem.byte_decode(B, d) == F).all()
    raw = rng.integers(0, 256, (n, 32 * d), dtype=np.uint8)
    assert (mlkem.byte_decode(raw, d) == oracle.byte_decode(raw, d)).all()
    x = rng.integers(0, 3329, (n, 256), dtype=np.uint16)
    cx = oracle.compress(x.ravel(), d).reshape(n, 256)
    assert (mlkem.compress(x, d) == cx).all()
    assert (mlkem.decompress(cx, d) == oracle.decompress(cx.ravel(), d).reshape(n, 256)).all()
    assert (mlkem.compress_encode(x, d) == oracle.byte_encode(cx, d)).all()
    assert (mlkem.decode_decompress(raw, d) == oracle.decompress(oracle.byte_decode(raw, d).ravel(), d).reshape(n, 256)).all()


@pytest.mark.parametrize("which,length", [(0, 800), (0, 1184), (0, 1568), (1, 64), (2, 800), (2, 1120), (2, 1600), (0, 136), (2, 168)])
def test_hash_vs_oracle(mlkem, oracle, which, length):
    rng = np.random.default_rng(length)
    data = rng.integers(0, 256, (777, length), dtype=np.uint8)
    assert (mlkem.hash_batch(which, data, length) == oracle.hash_batch(which, data, length)).all()


# ------------------------------------------------------------------ KEM vs oracle
def tamper(c, rng_seed=0):
    """BASELINE config 4: item i is tampered iff i % 10 == 3: flip bit (i % 8) of byte (i * 7919) % len."""
    c = c.copy()
    n, L = c.shape
    idx = np.arange(n)
    sel = idx[idx % 10 == 3]
    c[sel, (sel * 7919) % L] ^= (1 << (sel % 8)).astype(np.uint8)
    return c, sel


@pytest.mark.parametrize("ps", SETS)
@pytest.mark.parametrize("n", [1, 33, 2500])
def test_kem_vs_oracle(mlkem, oracle, ps, n):
    rng = np.random.default_rng(ps + n)
    d, z, m = (rng.integers(0, 256, (n, 32), dtype=np.uint8) for _ in range(3))
    ek, dk = mlkem.keygen(ps, d, z)
    oek, odk = oracle.keygen(ps, d, z)
    assert (ek == oek).all() and (dk == odk).all()
    c, K = mlkem.encaps(ps, ek, m)
    oc, oK = oracle.encaps(ps, ek, m)
    assert (c == oc).all() and (K == oK).all()
    bad, sel = tamper(c)
    Kd = mlkem.decaps(ps, dk, bad)
    assert (Kd == oracle.decaps(ps, dk, bad)).all()
    ok = np.ones(n, bool)
    ok[sel] = False
    assert (Kd[ok] == K[ok]).all()            # untampered: decapsulated key == encapsulated key
    assert (Kd[~ok] != K[~ok]).any(axis=1).all()  # tampered: implicit rejection
    assert (mlkem.check_dk(ps, dk) == 0).all()
    # K-PKE entry points
    pek, pdk = mlkem.pke_keygen(ps, d)
    assert (pek == ek).all() and (pdk == dk[:, : pdk.shape[1]]).all()
    r = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    cp = mlkem.pke_encrypt(ps, ek, m, r)
    assert (cp == oracle.pke_encrypt(ps, ek, m, r)).all()
    assert (mlkem.pke_decrypt(ps, pdk, cp) == m).all()
    assert (mlkem.pke_decrypt(ps, dk, cp, dk_stride=dk.shape[1]) == m).all()


@pytest.mark.parametrize("ps", SETS)
def test_kem_with_lowered_group_limit(oracle, ps):
    """The fused matrix kernel samples exactly three XOF blocks (168 groups) and leaves incomplete rows to the clean-up
    pass, which runs the general sampler with the give-up / restart rule of ml_kem.c:221-242.  Limits below 168 send
    every row through the general kernel; 168 and 170 keep the fused kernel and make its deferred rows restart."""
    import crystals_kyber_b200 as ck

    n = 700
    rng = np.random.default_rng(ps + 77)
    d, z, m = (rng.integers(0, 256, (n, 32), dtype=np.uint8) for _ in range(3))
    for limit in (158, 168, 170):  # 158: about half of the sponges restart (a much lower limit could cycle through all 256 seeds)
        gpu = ck.MLKEM(sample_group_limit=limit)
        oracle.set_sample_group_limit(limit)
        try:
            ek, dk = gpu.keygen(ps, d, z)
            oek, odk = oracle.keygen(ps, d, z)
            assert (ek == oek).all() and (dk == odk).all()
            c, K = gpu.encaps(ps, ek, m)
            oc, oK = oracle.encaps(ps, ek, m)
            assert (c == oc).all() and (K == oK).all()
            bad, sel = tamper(c)
            Kd = gpu.decaps(ps, dk, bad)
            assert (Kd == oracle.decaps(ps, dk, bad)).all()
            ok = np.ones(n, bool)
            ok[sel] = False
            assert (Kd[ok] == K[ok]).all() and (Kd[~ok] != K[~ok]).any(axis=1).all()
        finally:
            oracle.set_sample_group_limit(0)
    ek_ref, _ = oracle.keygen(ps, d, z)
    assert (ek_ref != oek).any(), "the lowered limit must change some matrix entries for the test to mean anything"


def test_kem_chunked_host_pipeline(oracle):
    """Host-memory path with several chunks alternating between the two pipeline slots."""
    import crystals_kyber_b200 as ck

    gpu = ck.MLKEM(chunk_items=96)
    rng = np.random.default_rng(11)
    n = 1000
    d, z, m = (rng.integers(0, 256, (n, 32), dtype=np.uint8) for _ in range(3))
    ek, dk = gpu.keygen(768, d, z)
    oek, odk = oracle.keygen(768, d, z)
    assert (ek == oek).all() and (dk == odk).all()
    c, K = gpu.encaps(768, ek, m)
    oc, oK = oracle.encaps(768, ek, m)
    assert (c == oc).all() and (K == oK).all()
    bad, _ = tamper(c)
    assert (gpu.decaps(768, dk, bad) == oracle.decaps(768, dk, bad)).all()


def test_check_dk_detects_corruption(mlkem):
    rng = np.random.default_rng(3)
    d, z = (rng.integers(0, 256, (64, 32), dtype=np.uint8) for _ in range(2))
    _, dk = mlkem.keygen(768, d, z)
    dk = dk.copy()
    dk[5, 1152 + 100] ^= 4   # inside the embedded ek
    dk[9, 2400 - 64 + 3] ^= 1  # inside the stored hash
    st = mlkem.check_dk(768, dk)
    assert st[5] == -5 and st[9] == -5 and (np.delete(st, [5, 9]) == 0).all()


def test_public_wrapper_batches(mlkem, oracle):
    """Batched KEM_KeyGen / KEM_Encaps / KEM_Decaps (ml_kem.c:1233-1359): entropy from the host source, the
    reference's length checks (-3) and per-item hash check (-5)."""
    n = 300
    ek, dk = mlkem.kem_keygen(768, n)
    assert len({bytes(r) for r in ek[:, -32:]}) == n  # fresh rho per item
    oek_check = [oracle.check_decaps_input(768, dk[i].tobytes(), 1088) for i in (0, 7, n - 1)]
    assert oek_check == [0, 0, 0]
    rc, c, K = mlkem.kem_encaps(768, ek)
    assert rc == 0
    bad_dk = dk.copy()
    bad_dk[17, 1152 + 5] ^= 1  # corrupt the embedded ek of item 17 -> hash check fails for that item only
    rc, Kd, status = mlkem.kem_decaps(768, bad_dk, c)
    assert rc == 0 and status[17] == -5 and (np.delete(status, 17) == 0).all()
    assert (Kd[17] == 0).all() and (np.delete(Kd, 17, 0) == np.delete(K, 17, 0)).all()
    assert (oracle.decaps(768, dk, c) == K).all()
    assert mlkem.kem_encaps(768, ek, ek_len=1)[0] == -3            # EncapsDecaps_test.c passes ek_len = 1
    assert mlkem.kem_decaps(768, dk, c, c_len=5)[0] == -3
    assert mlkem.kem_decaps(768, dk, c, dk_len=2399)[0] == -3
    assert mlkem.kem_encaps(999, ek)[0] == -1


@pytest.mark.parametrize("ps", SETS)
def test_fips_mode(oracle, ps):
    """MLKEM_B200_FLAG_FIPS203: SHAKE256 PRF / J and a real modulus check.  Checked against the oracle's FIPS switch
    (itself pinned to hashlib and to the `cryptography` package in tests/test_oracle.py) and, for 768 / 1024, directly
    against `cryptography`."""
    import crystals_kyber_b200 as ck

    gpu = ck.MLKEM(fips203=True)
    rng = np.random.default_rng(ps + 1)
    n = 700
    d, z, m = (rng.integers(0, 256, (n, 32), dtype=np.uint8) for _ in range(3))
    ek, dk = gpu.keygen(ps, d, z)
    c, K = gpu.encaps(ps, ek, m)
    bad, sel = tamper(c)
    Kd = gpu.decaps(ps, dk, bad)
    oracle.set_fips(True)
    try:
        oek, odk = oracle.keygen(ps, d, z)
        oc, oK = oracle.encaps(ps, oek, m)
        oKd = oracle.decaps(ps, odk, bad)
        s32 = rng.integers(0, 256, (n, 32), dtype=np.uint8)
        nonces = rng.integers(0, 256, n, dtype=np.uint8)
        for eta in (2, 3):
            assert (gpu.prf_cbd(s32, nonces, eta) == oracle.prf_cbd(s32, nonces, eta)).all()
    finally:
        oracle.set_fips(False)
    assert (ek == oek).all() and (dk == odk).all() and (c == oc).all() and (K == oK).all() and (Kd == oKd).all()
    ref_ek, _ = oracle.keygen(ps, d[:4], z[:4])
    assert (ref_ek != ek[:4]).any()  # and it really is a different function from the reference's
    data = rng.integers(0, 256, (50, 1120), dtype=np.uint8)
    assert [bytes(r) for r in gpu.hash_batch(3, data, 1120)] == [hashlib.shake_256(bytes(r)).digest(32) for r in data]
    # modulus check of KEM_Encaps: cannot fail in the reference (D4), fails with -4 here
    ek_bad = ek[:3].copy()
    ek_bad[1, 0:2] = 0xFF  # coefficient 0 of item 1 = 0xFFF >= q
    assert gpu.kem_encaps(ps, ek_bad)[0] == -4
    assert gpu.kem_encaps(ps, ek[:3])[0] == 0
    assert ck.MLKEM().kem_encaps(ps, ek_bad)[0] == 0
    if ps in (768, 1024):
        mlkem_mod = pytest.importorskip("cryptography.hazmat.primitives.asymmetric.mlkem")
        Priv = mlkem_mod.MLKEM768PrivateKey if ps == 768 else mlkem_mod.MLKEM1024PrivateKey
        for i in range(8):
            key = Priv.from_seed_bytes(d[i].tobytes() + z[i].tobytes())
            assert key.public_key().public_bytes_raw() == ek[i].tobytes()
            assert key.decapsulate(c[i].tobytes()) == K[i].tobytes()
            assert key.decapsulate(bad[i].tobytes()) == Kd[i].tobytes()


def test_sha3_front_end(mlkem, oracle, sha_examples):
    """sha3_b / sha3_s (sha3.c:408,465) on the GPU: the 16 NIST example files the reference's sha_testing.sh checks,
    reference outputs for lengths around the block boundary (incl. its padding deviation), random batches vs the oracle,
    and the reference-signature sha3_s through ctypes."""
    import ctypes as C

    for ex in sha_examples["examples"]:
        got = mlkem.sha3_bits([int(b) for b in ex["bits"]] if ex["bits"] else np.zeros((1, 0), np.uint8), ex["sfx"], ex["c"], ex["d"])[0]
        assert np.packbits(got, bitorder="little").tobytes().hex() == ex["hex"], ex["name"]
    for b in sha_examples["boundary"]:
        got = mlkem.sha3_bits([int(x) for x in b["bits"]], b["sfx"], b["c"], b["d"])[0]
        assert hashlib.sha256(got.tobytes()).hexdigest() == b["out_bits_sha256"]
    rng = np.random.default_rng(8)
    for sfx, c, d, nbits in (([0, 1, 0, 0], 512, 256, 1085), ([1, 1, 1, 1], 256, 3000, 2689), ([0, 1, 0, 0], 1024, 512, 7)):
        msgs = rng.integers(0, 2, (300, nbits), dtype=np.uint8)
        got = mlkem.sha3_bits(msgs, sfx, c, d)
        for i in (0, 1, 150, 299):
            assert (got[i] == oracle.sha3_bits(msgs[i], sfx, c, d)).all()
    lib = mlkem.lib
    lib.sha3_s.restype, lib.sha3_s.argtypes = C.POINTER(C.c_ubyte), [C.c_char_p, C.c_uint, C.c_uint, C.c_uint, C.POINTER(C.c_uint)]
    for rec in sha_examples["sha3_s"]:
        sfx = (C.c_uint * 4)(0, 1, 0, 0)
        t = rec["text"].encode()
        out = lib.sha3_s(t, len(t), 256, 512, sfx)
        assert bytes(out[i] for i in range(32)).hex() == rec["hex"] == hashlib.sha3_256(t).hexdigest()
        C.CDLL(None).free(out)


def test_device_memory_path(mlkem, oracle):
    """torch CUDA tensors in, torch CUDA tensors out, kernels on the current torch stream."""
    import torch

    rng = np.random.default_rng(21)
    n = 4096
    d, z, m = (rng.integers(0, 256, (n, 32), dtype=np.uint8) for _ in range(3))
    td, tz, tm = (torch.from_numpy(x).cuda() for x in (d, z, m))
    ek, dk = mlkem.keygen(768, td, tz)
    c, K = mlkem.encaps(768, ek, tm)
    K2 = mlkem.decaps(768, dk, c)
    torch.cuda.synchronize()
    oek, odk = oracle.keygen(768, d, z)
    oc, oK = oracle.encaps(768, oek, m)
    assert (ek.cpu().numpy() == oek).all() and (dk.cpu().numpy() == odk).all()
    assert (c.cpu().numpy() == oc).all() and (K.cpu().numpy() == oK).all() and (K2.cpu().numpy() == oK).all()
    f = torch.from_numpy(rng_polys(rng, 1024)).cuda()
    assert torch.equal(mlkem.intt(mlkem.ntt(f)).view(torch.int16).cpu(), f.view(torch.int16).cpu())


def test_full_size_properties(mlkem, oracle):
    """BASELINE sizes through size-independent properties: 2^20 polynomials round-trip through NTT/INTT,
    multiplication agrees with the oracle on a slice, and 2^18 encapsulations decapsulate to the same key
    (tampered ones are rejected), all on device memory."""
    import torch

    n = 1 << 20
    g = torch.Generator(device="cuda").manual_seed(20261018)
    f = torch.randint(0, 3329, (n, 256), generator=g, device="cuda", dtype=torch.int16).view(torch.uint16)
    h = torch.randint(0, 3329, (n, 256), generator=g, device="cuda", dtype=torch.int16).view(torch.uint16)
    fh = mlkem.ntt(f)
    back = mlkem.intt(fh)
    assert torch.equal(back.view(torch.int16), f.view(torch.int16))
    prod = mlkem.multiply_ntts(fh, mlkem.ntt(h))
    sl = slice(777_000, 777_000 + 4096)
    fs, hs = f[sl].cpu().numpy(), h[sl].cpu().numpy()
    assert (fh[sl].cpu().numpy() == oracle.ntt(fs)).all()
    assert (prod[sl].cpu().numpy() == oracle.multiply_ntts(oracle.ntt(fs), oracle.ntt(hs))).all()
    # linearity of the transform: NTT(f + h) == NTT(f) + NTT(h) (mod q)
    # (torch's uint16 support is storage-only: do the arithmetic on int16 views, all values are < 2^15)
    i32 = lambda t: t.view(torch.int16).to(torch.int32)
    s = ((i32(f) + i32(h)) % 3329).to(torch.int16).view(torch.uint16)
    lhs = i32(mlkem.ntt(s))
    rhs = (i32(fh) + i32(mlkem.ntt(h))) % 3329
    assert torch.equal(lhs, rhs)
    del f, h, fh, back, prod, s, lhs, rhs
    n = 1 << 18
    seeds = torch.randint(0, 256, (3, n, 32), generator=g, device="cuda", dtype=torch.uint8)
    ek, dk = mlkem.keygen(768, seeds[0], seeds[1])
    c, K = mlkem.encaps(768, ek, seeds[2])
    idx = torch.arange(n, device="cuda")
    sel = idx[idx % 10 == 3]
    c[sel, (sel * 7919) % c.shape[1]] ^= (1 << (sel % 8)).to(torch.uint8)
    Kd = mlkem.decaps(768, dk, c)
    same = (Kd == K).all(dim=1)
    assert bool(same[idx % 10 != 3].all()) and not bool(same[sel].any())
    sl = slice(100_000, 100_000 + 512)
    assert (Kd[sl].cpu().numpy() == oracle.decaps(768, dk[sl].cpu().numpy(), c[sl].cpu().numpy())).all()


@pytest.mark.parametrize("ps", SETS)
def test_baseline_config3_keygen_full_size(mlkem, oracle, ps):
    """BASELINE configs[2]: batched KeyGen for 2^20 keys per parameter set, seeds from the global index (workload.py).
    Size-independent properties over the whole batch -- dk embeds dk_pke || ek || H(ek) || z (ml_kem.c:1050-1077), so
    the embedded copy must equal ek, the device hash check (ml_kem.c:1336-1350) must pass for every key and fail when
    a byte of the embedded ek is flipped, z must be in place -- plus bit-exactness against the oracle on slices."""
    import torch

    import crystals_kyber_b200 as ck
    from crystals_kyber_b200 import workload as wl

    n, k = 1 << 20, ck.PARAMS[ps][0]
    d, z, _ = wl.derive_inputs(lambda msg, ln: mlkem.hash_batch(1, msg, ln), 0, n, torch.device("cuda"))
    ek, dk = mlkem.keygen(ps, d, z)
    assert torch.equal(dk[:, 384 * k : 768 * k + 32], ek)
    assert torch.equal(dk[:, 768 * k + 64 :], z)
    status = mlkem.check_dk(ps, dk)
    assert int((status != 0).sum()) == 0
    bad = dk[:4096].clone()
    bad[:, 384 * k + 5] ^= 1
    assert bool((mlkem.check_dk(ps, bad) == -5).all())
    for lo in (0, 500_000, n - 257):
        sl = slice(lo, lo + 257)
        oek, odk = oracle.keygen(ps, d[sl].cpu().numpy(), z[sl].cpu().numpy())
        assert (ek[sl].cpu().numpy() == oek).all() and (dk[sl].cpu().numpy() == odk).all()


def test_baseline_config4_encaps_decaps_full_size(mlkem, oracle):
    """BASELINE configs[3]: ML-KEM-768 Encaps + Decaps over 2^22 items, 10 % tampered (i % 10 == 3).  Properties over the
    whole batch: every untampered item decapsulates to the encapsulated key, every tampered one does not; slices
    (tampered items included) are bit-exact against the oracle, i.e. the rejection key is J(z || c') (ml_kem.c:1196-1215)."""
    import torch

    from crystals_kyber_b200 import workload as wl

    n = 1 << 22
    d, z, m = wl.derive_inputs(lambda msg, ln: mlkem.hash_batch(1, msg, ln), 0, n, torch.device("cuda"))
    ek, dk = mlkem.keygen(768, d, z)
    c, K = mlkem.encaps(768, ek, m)
    sel = wl.tamper_inplace(c, 0)
    Kd = mlkem.decaps(768, dk, c)
    same = (Kd == K).all(dim=1)
    tampered = torch.zeros(n, dtype=torch.bool, device="cuda")
    tampered[sel] = True
    assert int(sel.numel()) == (n + 6) // 10
    assert bool(same[~tampered].all()) and not bool(same[tampered].any())
    for lo in (0, 2_000_003, n - 300):
        sl = slice(lo, lo + 300)
        oc, oK = oracle.encaps(768, ek[sl].cpu().numpy(), m[sl].cpu().numpy())
        assert (K[sl].cpu().numpy() == oK).all()
        ct = c[sl].cpu().numpy()
        loc = (np.arange(lo, lo + 300) % 10 == 3)
        assert (ct[~loc] == oc[~loc]).all() and (ct[loc] != oc[loc]).any(axis=1).all()
        assert (Kd[sl].cpu().numpy() == oracle.decaps(768, dk[sl].cpu().numpy(), ct)).all()


# ------------------------------------------------------------------ the reference-signature API (include/ml_kem.h)
def test_reference_signature_api(mlkem, oracle):
    """KEM_KeyGen -> KEM_Encaps -> KEM_Decaps through the drop-in C API with the reference's stride-4 unions
    (BASELINE config 1 / EncapsDecaps_test.c with a correct ek_len), plus its error codes."""
    import ctypes as C

    lib = mlkem.lib

    class PARAMS(C.Structure):
        _fields_ = [("k", C.c_uint), ("n1", C.c_uint), ("n2", C.c_uint), ("du", C.c_uint), ("dv", C.c_uint)]

    class PKE(C.Structure):
        _fields_ = [("ek", C.POINTER(C.c_uint)), ("dk", C.POINTER(C.c_uint)), ("ek_len", C.c_uint), ("dk_len", C.c_uint)]

    class KEM(C.Structure):
        _fields_ = [("K", C.c_uint * 32), ("c", C.POINTER(C.c_uint)), ("c_len", C.c_uint)]

    assert C.sizeof(PARAMS) == 20 and C.sizeof(PKE) == 24 and C.sizeof(KEM) == 144  # D5
    lib.init.restype, lib.init.argtypes = PARAMS, [C.c_int]
    lib.KEM_KeyGen.restype, lib.KEM_KeyGen.argtypes = PKE, [C.POINTER(PARAMS)]
    lib.KEM_Encaps.restype, lib.KEM_Encaps.argtypes = KEM, [C.POINTER(PARAMS), C.POINTER(C.c_uint), C.c_uint]
    lib.KEM_Decaps.restype = C.POINTER(C.c_uint)
    lib.KEM_Decaps.argtypes = [C.POINTER(PARAMS), C.POINTER(C.c_uint), C.c_uint, C.POINTER(C.c_uint), C.c_uint]
    errno = C.c_int.in_dll(lib, "ml_errno")
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]

    p = lib.init(768)
    assert (p.k & 0xFF, p.n1 & 0xFF, p.n2 & 0xFF, p.du & 0xFF, p.dv & 0xFF) == (3, 2, 2, 10, 4)
    keys = lib.KEM_KeyGen(C.byref(p))
    assert keys.ek_len == 1184 and keys.dk_len == 2400 and errno.value == 0
    enc = lib.KEM_Encaps(C.byref(p), keys.ek, keys.ek_len)
    assert enc.c_len == 1088 and errno.value == 0
    Kd = lib.KEM_Decaps(C.byref(p), keys.dk, keys.dk_len, enc.c, enc.c_len)
    assert bool(Kd) and errno.value == 0
    K1 = bytes(enc.K[i] & 0xFF for i in range(32))
    K2 = bytes(Kd[i] & 0xFF for i in range(32))
    assert K1 == K2
    # the oracle agrees with what the drop-in API produced
    dk = np.array([keys.dk[i] & 0xFF for i in range(2400)], np.uint8)
    c = np.array([enc.c[i] & 0xFF for i in range(1088)], np.uint8)
    assert oracle.decaps(768, dk, c).tobytes() == K1
    # error behaviour (ml_kem.c:1267-1350): wrong ek_len -> -3 (what EncapsDecaps_test.c exercises), bad hash -> -5
    lib.KEM_Encaps(C.byref(p), keys.ek, 1)
    assert errno.value == -3
    errno.value = 0
    assert not lib.KEM_Decaps(C.byref(p), keys.dk, keys.dk_len, enc.c, 5) and errno.value == -3
    errno.value = 0
    keys.dk[2400 - 40] ^= 1
    assert not lib.KEM_Decaps(C.byref(p), keys.dk, keys.dk_len, enc.c, enc.c_len) and errno.value == -5
    errno.value = 0
    lib.init(999)
    assert errno.value == -1
    errno.value = 0
    for ptr in (keys.ek, keys.dk, enc.c, Kd):
        libc.free(ptr)


def test_second_device_from_one_process(mlkem, oracle):
    """opts.device selects the GPU: one process driving two devices (shard by index, no collective)."""
    import threading

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import crystals_kyber_b200 as ck
    from crystals_kyber_b200 import workload as wl

    n = 4000
    results = {}

    def run(dev):
        b, e = ck.shard_range(n, dev, 2)
        device = torch.device("cuda", dev)
        d, z, m = wl.derive_inputs(lambda msg, ln: mlkem.hash_batch(1, msg, ln), b, e, device)
        ek, dk = mlkem.keygen(768, d, z)
        c, K = mlkem.encaps(768, ek, m)
        Kd = mlkem.decaps(768, dk, c)
        torch.cuda.synchronize(device)
        results[dev] = tuple(t.cpu().numpy() for t in (ek, c, K, Kd))

    threads = [threading.Thread(target=run, args=(dev,)) for dev in (0, 1)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    d, z, m = wl.derive_inputs(lambda msg, ln: oracle.hash_batch(1, msg, ln), 0, n)
    oek, odk = oracle.keygen(768, d, z)
    oc, oK = oracle.encaps(768, oek, m)
    for dev in (0, 1):
        b, e = ck.shard_range(n, dev, 2)
        ek, c, K, Kd = results[dev]
        assert (ek == oek[b:e]).all() and (c == oc[b:e]).all() and (K == oK[b:e]).all() and (Kd == oK[b:e]).all()


def test_argument_errors_and_empty_batches(mlkem):
    import ctypes as C

    import torch

    from crystals_kyber_b200.lib import MEM_DEVICE, Opts

    lib = mlkem.lib
    buf = np.zeros(4096, np.uint8)
    p = C.c_void_p(buf.ctypes.data)
    assert lib.mlkem_b200_keygen_batch(768, 0, p, p, p, p, None) == 0           # empty batch: nothing to do
    assert lib.mlkem_b200_keygen_batch(999, 1, p, p, p, p, None) == -1          # ml_errno -1: unknown parameter set
    assert lib.mlkem_b200_keygen_batch(768, 1, None, p, p, p, None) == -11      # NULL buffer
    assert lib.mlkem_b200_byte_encode_batch(7, 1, p, p, None) == -11            # d not in {1,4,5,10,11,12}
    assert lib.mlkem_b200_cbd_batch(4, 1, p, p, None) == -11                    # eta not in {2,3}
    assert lib.mlkem_b200_compress_batch(10, 12, p, p, None) == -11             # element count must be a multiple of 8
    assert lib.mlkem_b200_hash_batch(0, 1, 13, p, p, None) == -11               # length must be a multiple of 8
    t = torch.zeros(8192, dtype=torch.uint8, device="cuda")
    o = Opts(0, MEM_DEVICE, None, 0, 0, 0)
    mis = C.c_void_p(t.data_ptr() + 4)
    ok = C.c_void_p(t.data_ptr())
    assert lib.mlkem_b200_ntt_batch(1, mis, ok, C.byref(o)) == -11              # device pointers must be 16-byte aligned
    assert b"aligned" in lib.mlkem_b200_last_error()
    assert mlkem.ntt(np.zeros((0, 256), np.uint16)).shape == (0, 256)
    assert mlkem.keygen(768, np.zeros((0, 32), np.uint8), np.zeros((0, 32), np.uint8))[0].shape == (0, 1184)


def test_unmodified_reference_drivers_linked_against_the_library(mlkem, archive_stdout):
    """The reference's own Test_Archive drivers, compiled UNMODIFIED from /root/reference against include/ml_kem.h and
    linked with libmlkem_b200.so instead of ml_kem.o (oracle/Makefile `drivers`; the binaries travel in oracle/_ref).
    Their stdout must be byte-identical to what they print when linked with the reference itself."""
    import os
    import subprocess

    ddir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "drivers")
    if not os.path.isdir(ddir):
        pytest.skip("oracle/_ref/drivers not built (needs /root/reference at build time)")
    ran = 0
    for name, rec in archive_stdout.items():
        exe = os.path.join(ddir, name)
        if not os.path.exists(exe):
            continue
        out = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
        assert hashlib.sha256(out.stdout).hexdigest() == rec["sha256"], (name, out.stdout[:200], out.stderr[:200])
        ran += 1
    assert ran >= 9
    # EncapsDecaps_test.c (makefile target test12) passes ek_len = 1: the reference prints the type-check error and exits 1
    out = subprocess.run([os.path.join(ddir, "EncapsDecaps_test")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    assert out.returncode == 1 and b"KEM_Encaps() :: Type check failed" in out.stderr and out.stdout == b""
    # KeyGen_test.c (target test11): random keys; structure only
    out = subprocess.run([os.path.join(ddir, "KeyGen_test")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    assert out.returncode == 0
    ek_line, dk_line = [l for l in out.stdout.decode().split("\n") if l.startswith(("Encapsulation key:", "Decapsulation key:"))]
    ek = [int(x) for x in ek_line.split(":")[1].split()]
    dk = [int(x) for x in dk_line.split(":")[1].split()]
    assert len(ek) == 800 and len(dk) == 1632 and dk[768:768 + 800] == ek
    assert bytes(dk[768 + 800:768 + 832]) == hashlib.sha3_256(bytes(ek)).digest()


def test_live_reference_if_present(mlkem, reference):
    """When oracle/_ref travelled to the box: the CUDA library against the compiled reference itself."""
    rng = np.random.default_rng(99)
    for ps in SETS:
        d, z, m = (rng.integers(0, 256, 32, dtype=np.uint8) for _ in range(3))
        ek, dk = mlkem.keygen(ps, d, z)
        rek, rdk = reference.keygen_internal(ps, d.tobytes(), z.tobytes())
        assert ek.tobytes() == rek and dk.tobytes() == rdk
        c, K = mlkem.encaps(ps, ek, m)
        rc, rK = reference.encaps_internal(ps, rek, m.tobytes())
        assert c.tobytes() == rc and K.tobytes() == rK
        bad = c.copy()
        bad[0, 17] ^= 0x20
        assert mlkem.decaps(ps, dk, bad).tobytes() == reference.decaps_internal(ps, rdk, bad.tobytes())
