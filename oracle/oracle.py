"""ctypes loaders for the CHECKERS under oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

`Oracle`   : oracle/libmlkem_oracle.so, the plain-C restatement (mlkem_oracle.c).
`Reference`: oracle/_ref/libref_mlkem.so, the unmodified reference compiled from
             /root/reference through ref_shim.c (only where it has been built).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module.  The product package (crystals-kyber_b200/) must never do so.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libmlkem_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libref_mlkem.so")
REF_G_SO = os.path.join(HERE, "_ref", "libref_mlkem_g.so")

PARAMS = {512: (2, 3, 2, 10, 4), 768: (3, 2, 2, 10, 4), 1024: (4, 2, 2, 11, 5)}


def sizes(param_set: int):
    k, _, _, du, dv = PARAMS[param_set]
    return {"k": k, "ek": 384 * k + 32, "dk": 768 * k + 96, "c": 32 * (du * k + dv), "dk_pke": 384 * k}


def build(force: bool = False) -> None:
    """Compile the checkers (make -C oracle).  Building the checker is not using it."""
    if force or not os.path.exists(ORACLE_SO) or (
        os.path.exists("/root/reference/ml_kem.c") and not os.path.exists(REF_SO)
    ):
        subprocess.run(["make", "-C", HERE], check=True, stdout=subprocess.DEVNULL)


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint8))


def _u16(a):
    a = np.ascontiguousarray(a, dtype=np.uint16)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint16))


class _Lib:
    def __init__(self, path, prefix):
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} not built (run `make -C oracle`)")
        self.lib = C.CDLL(path)
        self.p = prefix

    def fn(self, name, restype=None):
        f = getattr(self.lib, self.p + name)
        f.restype = restype
        return f


class Oracle(_Lib):
    """The plain-C restatement.  All arrays are numpy, dense, item-major."""

    def __init__(self, path=ORACLE_SO):
        super().__init__(path, "orc_")

    # ---- scalar / single-polynomial functions -------------------------------------------
    def bitrev7(self, r):
        return int(self.fn("bitrev7", C.c_uint8)(C.c_uint8(r)))

    def compress(self, x, d):
        x, px = _u16(np.atleast_1d(x))
        y = np.empty_like(x)
        self.fn("compress_batch")(C.c_size_t(x.size), C.c_uint(d), px, y.ctypes.data_as(C.POINTER(C.c_uint16)))
        return y

    def decompress(self, x, d):
        x, px = _u16(np.atleast_1d(x))
        y = np.empty_like(x)
        self.fn("decompress_batch")(C.c_size_t(x.size), C.c_uint(d), px, y.ctypes.data_as(C.POINTER(C.c_uint16)))
        return y

    def zeta_table(self):
        z = np.empty(128, np.uint16)
        self.fn("zeta_table")(z.ctypes.data_as(C.POINTER(C.c_uint16)))
        return z

    def gamma_table(self):
        z = np.empty(128, np.uint16)
        self.fn("gamma_table")(z.ctypes.data_as(C.POINTER(C.c_uint16)))
        return z

    def keccak_f1600(self, lanes):
        a = np.ascontiguousarray(lanes, dtype=np.uint64).copy()
        self.fn("keccak_f1600")(a.ctypes.data_as(C.POINTER(C.c_uint64)))
        return a

    def sponge(self, rate, dsfx, data: bytes, outlen):
        buf, p = _u8(np.frombuffer(data, np.uint8) if len(data) else np.zeros(0, np.uint8))
        out = np.empty(outlen, np.uint8)
        self.fn("sponge")(C.c_uint(rate), C.c_uint8(dsfx), p, C.c_size_t(len(data)),
                          out.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_size_t(outlen))
        return out.tobytes()

    def sha3_bits(self, msg_bits, sfx, c, d):
        """sha3_b on a list/array of bits; returns the d output bits as a numpy array."""
        bits = np.asarray(msg_bits, dtype=np.uint8)
        packed = np.packbits(bits, bitorder="little") if bits.size else np.zeros(1, np.uint8)
        out = np.zeros((d + 7) // 8, np.uint8)
        sf = np.asarray(sfx, dtype=np.uint8)
        rc = self.fn("sha3_bits", C.c_int)(packed.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_size_t(bits.size),
                                           sf.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_uint(c), C.c_size_t(d),
                                           out.ctypes.data_as(C.POINTER(C.c_uint8)))
        assert rc == 0
        return np.unpackbits(out, bitorder="little")[:d]

    def H(self, data: bytes):
        return self.sponge(136, 0x06, data, 32)

    def G(self, data: bytes):
        return self.sponge(72, 0x06, data, 64)

    def J(self, data: bytes):  # orc_J: SHAKE128 like the reference, SHAKE256 after set_fips(True)
        buf, p = _u8(np.frombuffer(data, np.uint8) if len(data) else np.zeros(0, np.uint8))
        out = np.empty(32, np.uint8)
        self.fn("J")(p, C.c_size_t(len(data)), out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out.tobytes()

    def PRF(self, s: bytes, b: int, eta: int):  # orc_PRF, same remark
        buf, p = _u8(np.frombuffer(s, np.uint8))
        out = np.empty(64 * eta, np.uint8)
        self.fn("PRF")(p, C.c_uint8(b), C.c_uint(eta), out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out.tobytes()

    # ---- batch functions ----------------------------------------------------------------
    def _polys(self, name, *ins):
        arrs = [_u16(a) for a in ins]
        n = arrs[0][0].size // 256
        out = np.empty((n, 256), np.uint16)
        self.fn(name)(C.c_size_t(n), *[p for _, p in arrs], out.ctypes.data_as(C.POINTER(C.c_uint16)))
        return out

    def ntt(self, f):
        return self._polys("ntt_batch", f)

    def intt(self, f):
        return self._polys("intt_batch", f)

    def multiply_ntts(self, f, g):
        return self._polys("multiply_ntts_batch", f, g)

    def poly_add(self, u, v):
        return self._polys("poly_add_batch", u, v)

    def poly_sub(self, u, v):
        return self._polys("poly_sub_batch", u, v)

    def vector_multiply(self, u, v, k):
        uu, pu = _u16(u)
        vv, pv = _u16(v)
        n = uu.size // (256 * k)
        out = np.empty((n, 256), np.uint16)
        self.fn("vector_multiply_batch")(C.c_size_t(n), C.c_uint(k), pu, pv, out.ctypes.data_as(C.POINTER(C.c_uint16)))
        return out

    def sample_ntt(self, seeds34):
        s, ps = _u8(seeds34)
        n = s.size // 34
        out = np.empty((n, 256), np.uint16)
        self.fn("sample_ntt_batch")(C.c_size_t(n), ps, out.ctypes.data_as(C.POINTER(C.c_uint16)))
        return out

    def sample_ntt_single(self, seed34):
        """Returns (coefficients, seed-after-call, restarts) -- B is in/out like ml_kem.c:189."""
        s = np.array(np.frombuffer(bytes(seed34), np.uint8))
        out = np.empty(256, np.uint16)
        r = self.fn("sample_ntt", C.c_int)(s.ctypes.data_as(C.POINTER(C.c_uint8)),
                                           out.ctypes.data_as(C.POINTER(C.c_uint16)))
        return out, s.tobytes(), int(r)

    def set_fips(self, on: bool):
        """FIPS 203 mode (SHAKE256 PRF/J, reducing ByteDecode12) -- not the reference's behaviour, see mlkem_oracle.c."""
        self.fn("set_fips")(C.c_int(1 if on else 0))

    def set_sample_group_limit(self, usable):
        """TEST HOOK, see orc_set_sample_group_limit (0 restores the reference's 278)."""
        self.fn("set_sample_group_limit")(C.c_uint(usable))

    def sample_ntt_with_seeds(self, seeds34):
        s = np.array(np.ascontiguousarray(seeds34, dtype=np.uint8)).reshape(-1, 34)
        out = np.empty((s.shape[0], 256), np.uint16)
        f = self.fn("sample_ntt", C.c_int)
        for i in range(s.shape[0]):
            f(s[i].ctypes.data_as(C.POINTER(C.c_uint8)), out[i].ctypes.data_as(C.POINTER(C.c_uint16)))
        return out, s

    def cbd(self, data, eta):
        b, pb = _u8(data)
        n = b.size // (64 * eta)
        out = np.empty((n, 256), np.uint16)
        self.fn("cbd_batch")(C.c_size_t(n), C.c_uint(eta), pb, out.ctypes.data_as(C.POINTER(C.c_uint16)))
        return out

    def prf_cbd(self, seeds32, nonces, eta):
        s, ps = _u8(seeds32)
        nn, pn = _u8(nonces)
        n = nn.size
        out = np.empty((n, 256), np.uint16)
        self.fn("prf_cbd_batch")(C.c_size_t(n), C.c_uint(eta), ps, pn, out.ctypes.data_as(C.POINTER(C.c_uint16)))
        return out

    def byte_encode(self, F, d):
        f, pf = _u16(F)
        n = f.size // 256
        out = np.empty((n, 32 * d), np.uint8)
        self.fn("byte_encode_batch")(C.c_size_t(n), C.c_uint(d), pf, out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out

    def byte_decode(self, B, d):
        b, pb = _u8(B)
        n = b.size // (32 * d)
        out = np.empty((n, 256), np.uint16)
        self.fn("byte_decode_batch")(C.c_size_t(n), C.c_uint(d), pb, out.ctypes.data_as(C.POINTER(C.c_uint16)))
        return out

    def hash_batch(self, which, data, length):
        b, pb = _u8(data)
        n = b.size // length if length else 0
        ol = 64 if which == 1 else 32
        out = np.empty((n, ol), np.uint8)
        self.fn("hash_batch")(C.c_int(which), C.c_size_t(n), C.c_size_t(length), pb,
                              out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out

    def keygen(self, ps, d, z):
        sz = sizes(ps)
        d, pd = _u8(d)
        z, pz = _u8(z)
        n = d.size // 32
        ek = np.empty((n, sz["ek"]), np.uint8)
        dk = np.empty((n, sz["dk"]), np.uint8)
        rc = self.fn("keygen_batch", C.c_int)(C.c_int(ps), C.c_size_t(n), pd, pz,
                                              ek.ctypes.data_as(C.POINTER(C.c_uint8)),
                                              dk.ctypes.data_as(C.POINTER(C.c_uint8)))
        assert rc == 0
        return ek, dk

    def encaps(self, ps, ek, m):
        sz = sizes(ps)
        ek, pe = _u8(ek)
        m, pm = _u8(m)
        n = m.size // 32
        c = np.empty((n, sz["c"]), np.uint8)
        K = np.empty((n, 32), np.uint8)
        rc = self.fn("encaps_batch", C.c_int)(C.c_int(ps), C.c_size_t(n), pe, pm,
                                              c.ctypes.data_as(C.POINTER(C.c_uint8)),
                                              K.ctypes.data_as(C.POINTER(C.c_uint8)))
        assert rc == 0
        return c, K

    def decaps(self, ps, dk, c):
        sz = sizes(ps)
        dk, pd = _u8(dk)
        c, pc = _u8(c)
        n = c.size // sz["c"]
        K = np.empty((n, 32), np.uint8)
        rc = self.fn("decaps_batch", C.c_int)(C.c_int(ps), C.c_size_t(n), pd, pc,
                                              K.ctypes.data_as(C.POINTER(C.c_uint8)))
        assert rc == 0
        return K

    def pke_keygen(self, ps, d):
        sz = sizes(ps)
        d, pd = _u8(d)
        n = d.size // 32
        ek = np.empty((n, sz["ek"]), np.uint8)
        dk = np.empty((n, sz["dk_pke"]), np.uint8)
        self.fn("pke_keygen_batch", C.c_int)(C.c_int(ps), C.c_size_t(n), pd,
                                             ek.ctypes.data_as(C.POINTER(C.c_uint8)),
                                             dk.ctypes.data_as(C.POINTER(C.c_uint8)))
        return ek, dk

    def pke_encrypt(self, ps, ek, m, r):
        sz = sizes(ps)
        ek, pe = _u8(ek)
        m, pm = _u8(m)
        r, pr = _u8(r)
        n = m.size // 32
        c = np.empty((n, sz["c"]), np.uint8)
        self.fn("pke_encrypt_batch", C.c_int)(C.c_int(ps), C.c_size_t(n), pe, pm, pr,
                                              c.ctypes.data_as(C.POINTER(C.c_uint8)))
        return c

    def pke_decrypt(self, ps, dk, c, dk_stride=None):
        sz = sizes(ps)
        dk, pd = _u8(dk)
        c, pc = _u8(c)
        n = c.size // sz["c"]
        stride = dk_stride if dk_stride is not None else dk.size // n
        m = np.empty((n, 32), np.uint8)
        self.fn("pke_decrypt_batch", C.c_int)(C.c_int(ps), C.c_size_t(n), C.c_size_t(stride), pd, pc,
                                              m.ctypes.data_as(C.POINTER(C.c_uint8)))
        return m

    def check_encaps_input(self, ps, ek: bytes):
        b, pb = _u8(np.frombuffer(ek, np.uint8))
        return int(self.fn("check_encaps_input", C.c_int)(C.c_int(ps), pb, C.c_uint(len(ek))))

    def check_decaps_input(self, ps, dk: bytes, c_len: int):
        b, pb = _u8(np.frombuffer(dk, np.uint8))
        return int(self.fn("check_decaps_input", C.c_int)(C.c_int(ps), pb, C.c_uint(len(dk)), C.c_uint(c_len)))


class Reference(_Lib):
    """The unmodified reference (ml_kem.c + sha3.c) behind dense wrappers.  Single items only."""

    def __init__(self, path=REF_SO):
        super().__init__(path, "ref_")

    def sizes(self):
        out = (C.c_uint * 6)()
        self.fn("sizes")(out)
        return list(out)

    def init(self, ps):
        out = (C.c_uint * 5)()
        rc = self.fn("init", C.c_int)(C.c_int(ps), out)
        return rc, tuple(out)

    def bitrev7(self, r):
        return int(self.fn("bitrev7", C.c_uint8)(C.c_uint8(r)))

    def compress(self, x, d):
        return int(self.fn("compress", C.c_uint16)(C.c_uint16(x), C.c_uint(d)))

    def decompress(self, x, d):
        return int(self.fn("decompress", C.c_uint16)(C.c_uint16(x), C.c_uint(d)))

    def byte_encode(self, F, d):
        f, pf = _u16(F)
        out = np.empty(32 * d, np.uint8)
        self.fn("byte_encode")(pf, C.c_uint(d), out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out

    def byte_decode(self, B, d):
        b, pb = _u8(B)
        out = np.empty(256, np.uint16)
        self.fn("byte_decode")(pb, C.c_uint(d), out.ctypes.data_as(C.POINTER(C.c_uint16)))
        return out

    def sample_ntt(self, seed34):
        s = np.array(np.frombuffer(bytes(seed34), np.uint8))
        out = np.empty(256, np.uint16)
        self.fn("sample_ntt")(s.ctypes.data_as(C.POINTER(C.c_uint8)), out.ctypes.data_as(C.POINTER(C.c_uint16)))
        return out, s.tobytes()

    def sample_cbd(self, data, eta):
        b, pb = _u8(np.frombuffer(bytes(data), np.uint8))
        out = np.empty(256, np.uint16)
        self.fn("sample_cbd")(pb, C.c_uint(eta), out.ctypes.data_as(C.POINTER(C.c_uint16)))
        return out

    def _poly(self, name, *ins):
        arrs = [_u16(a) for a in ins]
        out = np.empty(256, np.uint16)
        self.fn(name)(*[p for _, p in arrs], out.ctypes.data_as(C.POINTER(C.c_uint16)))
        return out

    def ntt(self, f):
        return self._poly("ntt", f)

    def intt(self, f):
        return self._poly("intt", f)

    def multiply_ntts(self, f, g):
        return self._poly("multiply_ntts", f, g)

    def poly_add(self, f, g):
        return self._poly("poly_add", f, g)

    def poly_sub(self, f, g):
        return self._poly("poly_sub", f, g)

    def vector_multiply(self, u, v, k):
        _, pu = _u16(u)
        _, pv = _u16(v)
        out = np.empty(256, np.uint16)
        self.fn("vector_multiply")(C.c_uint(k), pu, pv, out.ctypes.data_as(C.POINTER(C.c_uint16)))
        return out

    def basecase_multiply(self, a0, a1, b0, b1, gamma):
        out = (C.c_uint16 * 2)()
        self.fn("basecase_multiply")(C.c_uint16(a0), C.c_uint16(a1), C.c_uint16(b0), C.c_uint16(b1),
                                     C.c_uint16(gamma), out)
        return int(out[0]), int(out[1])

    def sha3_bits(self, msg_bits, sfx, c, d):
        bits = np.asarray(msg_bits, dtype=np.uint8)
        packed = np.packbits(bits, bitorder="little") if bits.size else np.zeros(1, np.uint8)
        out = np.zeros((d + 7) // 8, np.uint8)
        sf = np.asarray(sfx, dtype=np.uint8)
        self.fn("sha3_bits")(packed.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_uint(bits.size),
                             sf.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_uint(c), C.c_uint(d),
                             out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return np.unpackbits(out, bitorder="little")[:d]

    def sha3_s(self, text: bytes, sfx, c, d):
        out = np.zeros(d // 8, np.uint8)
        sf = np.asarray(sfx, dtype=np.uint8)
        self.fn("sha3_s")(C.c_char_p(text), C.c_uint(len(text)), sf.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_uint(c), C.c_uint(d),
                          out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out.tobytes()

    def _hash(self, name, data: bytes, outlen):
        b, pb = _u8(np.frombuffer(data, np.uint8) if len(data) else np.zeros(1, np.uint8))
        out = np.empty(outlen, np.uint8)
        self.fn(name)(pb, C.c_uint(len(data)), out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out.tobytes()

    def H(self, data):
        return self._hash("H", data, 32)

    def J(self, data):
        return self._hash("J", data, 32)

    def G(self, data):
        return self._hash("G", data, 64)

    def PRF(self, s: bytes, b: int, eta: int):
        sb, ps = _u8(np.frombuffer(s, np.uint8))
        out = np.empty(64 * eta, np.uint8)
        self.fn("PRF")(ps, C.c_uint8(b), C.c_uint(eta), out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out.tobytes()

    def pke_keygen(self, ps, d: bytes):
        sz = sizes(ps)
        _, pd = _u8(np.frombuffer(d, np.uint8))
        ek = np.empty(sz["ek"], np.uint8)
        dk = np.empty(sz["dk_pke"], np.uint8)
        self.fn("pke_keygen", C.c_int)(C.c_int(ps), pd, ek.ctypes.data_as(C.POINTER(C.c_uint8)),
                                       dk.ctypes.data_as(C.POINTER(C.c_uint8)))
        return ek.tobytes(), dk.tobytes()

    def pke_encrypt(self, ps, ek: bytes, m: bytes, r: bytes):
        sz = sizes(ps)
        ins = [_u8(np.frombuffer(x, np.uint8)) for x in (ek, m, r)]
        c = np.empty(sz["c"], np.uint8)
        self.fn("pke_encrypt", C.c_int)(C.c_int(ps), *[p for _, p in ins], c.ctypes.data_as(C.POINTER(C.c_uint8)))
        return c.tobytes()

    def pke_decrypt(self, ps, dk: bytes, c: bytes):
        ins = [_u8(np.frombuffer(x, np.uint8)) for x in (dk, c)]
        m = np.empty(32, np.uint8)
        self.fn("pke_decrypt", C.c_int)(C.c_int(ps), *[p for _, p in ins], m.ctypes.data_as(C.POINTER(C.c_uint8)))
        return m.tobytes()

    def keygen_internal(self, ps, d: bytes, z: bytes):
        sz = sizes(ps)
        ins = [_u8(np.frombuffer(x, np.uint8)) for x in (d, z)]
        ek = np.empty(sz["ek"], np.uint8)
        dk = np.empty(sz["dk"], np.uint8)
        self.fn("keygen_internal", C.c_int)(C.c_int(ps), *[p for _, p in ins],
                                            ek.ctypes.data_as(C.POINTER(C.c_uint8)),
                                            dk.ctypes.data_as(C.POINTER(C.c_uint8)))
        return ek.tobytes(), dk.tobytes()

    def encaps_internal(self, ps, ek: bytes, m: bytes):
        sz = sizes(ps)
        ins = [_u8(np.frombuffer(x, np.uint8)) for x in (ek, m)]
        c = np.empty(sz["c"], np.uint8)
        K = np.empty(32, np.uint8)
        self.fn("encaps_internal", C.c_int)(C.c_int(ps), *[p for _, p in ins],
                                            c.ctypes.data_as(C.POINTER(C.c_uint8)),
                                            K.ctypes.data_as(C.POINTER(C.c_uint8)))
        return c.tobytes(), K.tobytes()

    def decaps_internal(self, ps, dk: bytes, c: bytes):
        ins = [_u8(np.frombuffer(x, np.uint8)) for x in (dk, c)]
        K = np.empty(32, np.uint8)
        self.fn("decaps_internal", C.c_int)(C.c_int(ps), *[p for _, p in ins],
                                            K.ctypes.data_as(C.POINTER(C.c_uint8)))
        return K.tobytes()

    def KEM_KeyGen(self, ps):
        sz = sizes(ps)
        ek = np.empty(sz["ek"], np.uint8)
        dk = np.empty(sz["dk"], np.uint8)
        rc = self.fn("KEM_KeyGen", C.c_int)(C.c_int(ps), ek.ctypes.data_as(C.POINTER(C.c_uint8)),
                                            dk.ctypes.data_as(C.POINTER(C.c_uint8)))
        return rc, ek.tobytes(), dk.tobytes()

    def KEM_Encaps(self, ps, ek: bytes, ek_len=None):
        sz = sizes(ps)
        _, pe = _u8(np.frombuffer(ek, np.uint8))
        c = np.empty(sz["c"], np.uint8)
        K = np.empty(32, np.uint8)
        rc = self.fn("KEM_Encaps", C.c_int)(C.c_int(ps), pe, C.c_uint(len(ek) if ek_len is None else ek_len),
                                            c.ctypes.data_as(C.POINTER(C.c_uint8)),
                                            K.ctypes.data_as(C.POINTER(C.c_uint8)))
        return rc, c.tobytes(), K.tobytes()

    def KEM_Decaps(self, ps, dk: bytes, c: bytes, dk_len=None, c_len=None):
        _, pd = _u8(np.frombuffer(dk, np.uint8))
        _, pc = _u8(np.frombuffer(c, np.uint8))
        K = np.empty(32, np.uint8)
        rc = self.fn("KEM_Decaps", C.c_int)(C.c_int(ps), pd, C.c_uint(len(dk) if dk_len is None else dk_len), pc,
                                            C.c_uint(len(c) if c_len is None else c_len),
                                            K.ctypes.data_as(C.POINTER(C.c_uint8)))
        return rc, K.tobytes()

    # ---- timing drivers (bench.py) --------------------------------------------------------
    def time_pairs(self, ps, ek, dk, m, threads):
        sz = sizes(ps)
        ek, pe = _u8(ek)
        dk, pd = _u8(dk)
        m, pm = _u8(m)
        n = m.size // 32
        c = np.empty((n, sz["c"]), np.uint8)
        K = np.empty((n, 32), np.uint8)
        t = self.fn("time_pairs", C.c_double)(C.c_int(ps), C.c_size_t(n), C.c_int(threads), pe, pd, pm,
                                              c.ctypes.data_as(C.POINTER(C.c_uint8)),
                                              K.ctypes.data_as(C.POINTER(C.c_uint8)))
        return float(t), c, K

    def time_keygen(self, ps, d, z, threads):
        sz = sizes(ps)
        d, pd = _u8(d)
        z, pz = _u8(z)
        n = d.size // 32
        ek = np.empty((n, sz["ek"]), np.uint8)
        dk = np.empty((n, sz["dk"]), np.uint8)
        t = self.fn("time_keygen", C.c_double)(C.c_int(ps), C.c_size_t(n), C.c_int(threads), pd, pz,
                                               ek.ctypes.data_as(C.POINTER(C.c_uint8)),
                                               dk.ctypes.data_as(C.POINTER(C.c_uint8)))
        return float(t), ek, dk

    def time_encaps(self, ps, ek, m, threads):
        sz = sizes(ps)
        ek, pe = _u8(ek)
        m, pm = _u8(m)
        n = m.size // 32
        c = np.empty((n, sz["c"]), np.uint8)
        K = np.empty((n, 32), np.uint8)
        t = self.fn("time_encaps", C.c_double)(C.c_int(ps), C.c_size_t(n), C.c_int(threads), pe, pm,
                                               c.ctypes.data_as(C.POINTER(C.c_uint8)),
                                               K.ctypes.data_as(C.POINTER(C.c_uint8)))
        return float(t), c, K

    def time_decaps(self, ps, dk, c, threads):
        dk, pd = _u8(dk)
        c, pc = _u8(c)
        n = c.size // sizes(ps)["c"]
        K = np.empty((n, 32), np.uint8)
        t = self.fn("time_decaps", C.c_double)(C.c_int(ps), C.c_size_t(n), C.c_int(threads), pd, pc,
                                               K.ctypes.data_as(C.POINTER(C.c_uint8)))
        return float(t), K

    def time_ring(self, f, g, reps):
        _, pf = _u16(f)
        _, pg = _u16(g)
        out = (C.c_double * 3)()
        self.fn("time_ring")(C.c_size_t(reps), pf, pg, out)
        return tuple(out)
