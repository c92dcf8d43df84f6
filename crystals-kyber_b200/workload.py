"""Synthetic workloads of BASELINE.json (configs 3-5) -- seed derivation and the tamper rule.

Every per-item input is a function of the item's GLOBAL index only, so the concatenated outputs of a
sharded run do not depend on the number of GPUs:

    (d_i || z_i) = G("mlkemkey" || LE64(i))      64 bytes, G = SHA3-512
     m_i         = G("mlkemmsg" || LE64(i))[:32]

`hash_G` is the SHA3-512 batch function of whoever runs the workload: the CUDA library on the GPU
(`MLKEM.hash_batch(1, ...)`), the oracle in the CPU-only tests.

Tamper rule (config 4): ciphertext i is tampered iff i % 10 == 3, by flipping bit (i % 8) of byte
(i * 7919) % len(c).  Expected: untampered items decapsulate to the encapsulated key, tampered ones to
J(z || c') (implicit rejection).
"""
import numpy as np

TAG_KEY = b"mlkemkey"
TAG_MSG = b"mlkemmsg"


def _messages_np(tag: bytes, begin: int, end: int):
    idx = np.arange(begin, end, dtype="<u8")
    msg = np.empty((end - begin, 16), np.uint8)
    msg[:, :8] = np.frombuffer(tag, np.uint8)
    msg[:, 8:] = idx.view(np.uint8).reshape(-1, 8)
    return msg


def _messages_torch(tag: bytes, begin: int, end: int, device):
    import torch

    idx = torch.arange(begin, end, dtype=torch.int64, device=device)
    msg = torch.empty((end - begin, 16), dtype=torch.uint8, device=device)
    msg[:, :8] = torch.tensor(list(tag), dtype=torch.uint8, device=device)
    msg[:, 8:] = idx.view(torch.uint8).reshape(-1, 8)
    return msg


def derive_inputs(hash_G, begin: int, end: int, device=None):
    """Returns (d, z, m) for global items [begin, end).  device=None -> numpy, else torch tensors on device."""
    if device is None:
        key = hash_G(_messages_np(TAG_KEY, begin, end), 16)
        msg = hash_G(_messages_np(TAG_MSG, begin, end), 16)
        return np.ascontiguousarray(key[:, :32]), np.ascontiguousarray(key[:, 32:]), np.ascontiguousarray(msg[:, :32])
    key = hash_G(_messages_torch(TAG_KEY, begin, end, device), 16)
    msg = hash_G(_messages_torch(TAG_MSG, begin, end, device), 16)
    return key[:, :32].contiguous(), key[:, 32:].contiguous(), msg[:, :32].contiguous()


def tamper_inplace(c, begin: int):
    """Apply the tamper rule to ciphertexts of global items [begin, begin + len(c)).  Returns the local
    indices that were tampered.  Works on numpy arrays and torch tensors."""
    n, L = c.shape
    if isinstance(c, np.ndarray):
        g = np.arange(begin, begin + n, dtype=np.int64)
        sel = np.nonzero(g % 10 == 3)[0]
        gs = g[sel]
        c[sel, (gs * 7919) % L] ^= (1 << (gs % 8)).astype(np.uint8)
        return sel
    import torch

    g = torch.arange(begin, begin + n, dtype=torch.int64, device=c.device)
    sel = torch.nonzero(g % 10 == 3).flatten()
    gs = g[sel]
    c[sel, (gs * 7919) % L] ^= (1 << (gs % 8)).to(torch.uint8)
    return sel


# Checksum of checksums (SURVEY.md 8(e) "concatenated outputs at G = 1, 2, 4, 8 are byte-identical"): an output array is cut into
# blocks of `block` consecutive items, each block hashed, and the digest is the hash of the concatenated block hashes.  A rank
# whose shard starts and ends on block boundaries computes its part alone; concatenating the parts in rank order gives the digest
# of the whole range whatever the number of ranks.
def block_hashes(array, block: int) -> bytes:
    """sha256 of every `block`-item slice of a host array (numpy, item-major), concatenated."""
    import hashlib

    a = np.ascontiguousarray(array)
    assert a.shape[0] % block == 0, "a shard must be a whole number of blocks"
    return b"".join(hashlib.sha256(a[b : b + block].tobytes()).digest() for b in range(0, a.shape[0], block))


def combine_block_hashes(parts) -> str:
    """parts: the block_hashes() of the shards in rank order.  Returns the hex digest of the whole range."""
    import hashlib

    return hashlib.sha256(b"".join(parts)).hexdigest()


# Algorithmic work model, in 32-bit integer-pipe operations.
OPS_KECCAK_F = 24 * 180            # 4 320
OPS_NTT = 896 * 5                  # 4 480
OPS_INTT = 896 * 5 + 256 * 3       # 5 248
OPS_MULNTT = 128 * (5 * 3 + 2)     # 2 176


def keccak_calls(k: int, eta1: int, du: int, dv: int):
    """Keccak-f[1600] calls per operation at 3 squeeze blocks per SampleNTT (the expected 3.009)."""
    prf1 = 1 if eta1 == 2 else 2
    h = (384 * k + 32) // 136 + 1
    j = (32 + 32 * (du * k + dv)) // 168 + 1
    keygen = 1 + 3 * k * k + 2 * k * prf1 + h
    encrypt = 3 * k * k + k * prf1 + (k + 1)
    return {"keygen": keygen, "encaps": h + 1 + encrypt, "decaps": 1 + j + encrypt, "encrypt": encrypt}


def op_counts(k: int, eta1: int, du: int, dv: int):
    """Algorithmic INT32 operations per ML-KEM operation: Keccak + NTT + INTT + base multiplication."""
    kc = keccak_calls(k, eta1, du, dv)
    keygen = kc["keygen"] * OPS_KECCAK_F + 2 * k * OPS_NTT + k * k * OPS_MULNTT
    enc_arith = k * OPS_NTT + (k * k + k) * OPS_MULNTT + (k + 1) * OPS_INTT
    encaps = kc["encaps"] * OPS_KECCAK_F + enc_arith
    decaps = kc["decaps"] * OPS_KECCAK_F + enc_arith + k * OPS_NTT + k * OPS_MULNTT + OPS_INTT
    # the fused matrix-expansion + matrix-vector kernel of Encrypt: k^2 sponges, k^2 base multiplications, k inverse NTTs
    matvec_enc = 3 * k * k * OPS_KECCAK_F + k * k * OPS_MULNTT + k * OPS_INTT
    matvec_keygen = 3 * k * k * OPS_KECCAK_F + k * k * OPS_MULNTT
    return {"keygen": keygen, "encaps": encaps, "decaps": decaps, "matvec_encrypt": matvec_enc, "matvec_keygen": matvec_keygen}
