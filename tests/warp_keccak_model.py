"""numpy model of keccak_f1600_warp / warp_sponge_absorb_words (crystals-kyber_b200/csrc/mlkem_device.cuh): one state lane per
GPU lane, neighbours by shuffle (an index permutation of a 32-entry array here).  Pins the shuffle sources, the per-lane rho
offsets, the variable funnel-shift rotate and the pad placement against hashlib on the CPU (tests/test_abi_cpu.py)."""
import hashlib
import numpy as np

M32 = 0xFFFFFFFF
RHO = [0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14]
RC = []
lfsr = 1
for rnd in range(24):
    c = 0
    for j in range(7):
        if lfsr & 1:
            c |= 1 << ((1 << j) - 1)
        hi = lfsr & 0x80
        lfsr = (lfsr << 1) & 0xFF
        if hi:
            lfsr ^= 0x71
    RC.append(c)

lane = np.arange(32)
x, y = lane % 5, lane // 5
act = lane < 25
def idx(v):  # lanes >= 25 shuffle from themselves
    return np.where(act, v, lane)
s5, s10, s15, s20 = (idx((lane + k) % 25) for k in (5, 10, 15, 20))
xm, xp = idx((x + 4) % 5), idx((x + 1) % 5)
pi_src = idx((x + 3 * y) % 5 + 5 * x)
c1, c2 = idx(5 * y + (x + 1) % 5), idx(5 * y + (x + 2) % 5)
rot = np.array(RHO + [0] * 7)

def fsl(lo, hi, s):  # __funnelshift_l(lo, hi, s): upper 32 bits of (hi:lo) << (s & 31)
    v = (hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)
    return ((v << (s & 31).astype(np.uint64)) >> np.uint64(32)).astype(np.uint64) & np.uint64(M32)

def perm(lo, hi):
    lo, hi = lo.copy(), hi.copy()
    for rnd in range(24):
        clo = lo ^ lo[s5] ^ lo[s10] ^ lo[s15] ^ lo[s20]
        chi_ = hi ^ hi[s5] ^ hi[s10] ^ hi[s15] ^ hi[s20]
        mlo, mhi, plo, phi = clo[xm], chi_[xm], clo[xp], chi_[xp]
        rhi, rlo = fsl(plo, phi, np.full(32, 1)), fsl(phi, plo, np.full(32, 1))
        lo ^= mlo ^ rlo
        hi ^= mhi ^ rhi
        sw = (rot & 32) != 0
        l2, h2 = np.where(sw, hi, lo), np.where(sw, lo, hi)
        s = rot & 31
        hi, lo = fsl(l2, h2, s), fsl(h2, l2, s)
        blo, bhi = lo[pi_src], hi[pi_src]
        lo = blo ^ (~blo[c1] & blo[c2] & np.uint64(M32))
        hi = bhi ^ (~bhi[c1] & bhi[c2] & np.uint64(M32))
        lo[0] ^= np.uint64(RC[rnd] & M32)
        hi[0] ^= np.uint64(RC[rnd] >> 32)
        lo[25:] = 0  # (idle lanes: whatever they hold is never read by lanes < 25; kept zero here like in the kernel)
        hi[25:] = 0
    return lo, hi

def sponge(msg: bytes, rate_lanes: int, sfx: int, outlanes: int):
    assert len(msg) % 8 == 0
    words = np.frombuffer(msg, dtype="<u8")
    lo, hi = np.zeros(32, np.uint64), np.zeros(32, np.uint64)
    base = 0
    n = len(words)
    while base + rate_lanes <= n:
        w = np.zeros(32, np.uint64)
        w[:rate_lanes] = words[base:base + rate_lanes]
        lo ^= w & np.uint64(M32)
        hi ^= w >> np.uint64(32)
        lo, hi = perm(lo, hi)
        base += rate_lanes
    rem = n - base
    w = np.zeros(32, np.uint64)
    w[:rem] = words[base:]
    lo ^= w & np.uint64(M32)
    hi ^= w >> np.uint64(32)
    lo[rem] ^= np.uint64(sfx)
    hi[rate_lanes - 1] ^= np.uint64(0x80000000)
    lo, hi = perm(lo, hi)
    out = (hi[:outlanes] << np.uint64(32)) | lo[:outlanes]
    return out.astype("<u8").tobytes()

def check_against_hashlib():
    rng = np.random.default_rng(5)
    for ln in (1184, 800, 1568, 136, 128, 0, 8):
        m = rng.integers(0, 256, ln, dtype=np.uint8).tobytes()
        assert sponge(m, 17, 0x06, 4) == hashlib.sha3_256(m).digest(), ln
    for ln in (1120, 800, 1600, 168, 160):
        m = rng.integers(0, 256, ln, dtype=np.uint8).tobytes()
        assert sponge(m, 21, 0x1F, 4) == hashlib.shake_128(m).digest(32), ln
        assert sponge(m, 17, 0x1F, 4) == hashlib.shake_256(m).digest(32), ln
    m = rng.integers(0, 256, 64, dtype=np.uint8).tobytes()
    assert sponge(m, 9, 0x06, 8) == hashlib.sha3_512(m).digest()
    return True
