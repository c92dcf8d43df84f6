"""Re-statements of the reference's Test_Archive drivers as stdout generators.

Each function reproduces, byte for byte, what the corresponding reference driver prints
(/root/reference/Test_Archive/<name>.c) when given an implementation object `impl` that offers the
functions the driver calls.  tests compare sha256(stdout) with tests/golden/archive_stdout.json, which
was produced by building and running the reference's own drivers (tests/golden/make_golden.py).

`impl` is either the oracle (CPU tests) or the CUDA library's host mirror (GPU tests); both expose
bitrev7, byte_encode/byte_decode, compress/decompress, sample_ntt, cbd, ntt, intt, zeta_table,
pke_keygen, pke_encrypt, pke_decrypt with numpy arrays.
"""
import numpy as np


def bitrev7_test01(impl):  # BitRev7_test01.c:15-28
    out = []
    for i in range(128):
        out.append("Initial value: 0b" + format(i, "07b") + "\n")
        out.append("Reversed value: 0b" + format(int(impl.bitrev7(i)), "07b") + "\n\n")
    return "".join(out)


def bits_and_bytes_test02(impl):  # BitsAndBytes_test02.c: bits 01010101 -> 170 -> bits
    bits = [i % 2 for i in range(8)]
    F = np.zeros(256, np.uint16)
    F[:8] = bits
    byte0 = int(impl.byte_encode(F[None, :], 1)[0, 0])  # BitsToBytes == ByteEncode_1 on the first 8 bits
    back = impl.byte_decode(np.array([[byte0] + [0] * 31], np.uint8), 1)[0, :8]
    s = "Bits: " + "".join(str(b) for b in bits) + "\n"
    s += f"Bytes: {byte0}\n"
    s += "Bits: " + "".join(str(int(b)) for b in back) + "\n"
    return s


def encode_decode_test03(impl):  # EncodeDecode_test03.c:26-43
    F = (np.arange(256) * 16).astype(np.uint16)
    back = impl.byte_decode(impl.byte_encode(F[None, :], 12), 12)[0]
    fails = [i for i in range(256) if back[i] != F[i]]
    if not fails:
        return "Test Successful!\n"
    return "".join(f"Test Failed: {i}\n" for i in fails) + f"Fail Count: {len(fails)}\n"


def compress_decompress_test04(impl):  # CompressDecompress_test04.c:7-69
    out = []
    per_d = {}
    for d in range(1, 13):
        lim = 3329 if d == 12 else (1 << d)
        y = np.arange(lim, dtype=np.uint16)
        per_d[d] = (y, impl.compress(impl.decompress(y, d), d))
    for i in range(3329):
        for d in range(1, 13):
            y, x = per_d[d]
            if i < len(y) and x[i] != y[i]:
                out.append(f"ERROR-{d}:: y.t={y[i]} - x.t={x[i]}\n")
    out.append("Test Complete!\n")
    return "".join(out)


def _poly_str(a):
    return "".join(f"{int(c)}x^{i} + " for i, c in enumerate(a) if c != 0)


def sample_ntt_test06(impl):  # SampleNTT_test06.c:9-25
    seeds = np.array([[(it * i + i) & 0xFF for i in range(34)] for it in range(7)], np.uint8)
    a = impl.sample_ntt(seeds)
    return "".join(_poly_str(a[it]) for it in range(7))


def sample_cbd_test07(impl):  # SampleCBD_test07.c:8-19
    B = np.arange(192, dtype=np.uint8)
    return _poly_str(impl.cbd(B[None, :], 3)[0])


def ntt_test08(impl):  # NTT_test08.c:10-22
    seed = np.array([[(2 * i) & 0xFF for i in range(34)]], np.uint8)
    f1 = impl.sample_ntt(seed)
    f2 = impl.intt(impl.ntt(f1))
    out = [f"ERROR :: f1[{i}] = {f1[0, i]} :: f2[{i}] = {f2[0, i]}\n" for i in range(256) if f1[0, i] != f2[0, i]]
    return "".join(out) + "Test Complete!\n"


ZETA_TABLE_FIPS203 = None  # the driver hard-codes the FIPS 203 table; we compare against 17^BitRev7(i) directly


def zeta_logic_test(impl):  # ZetaLogic_test.c:21-37
    z = impl.zeta_table()
    out = []
    for i in range(128):
        expect = pow(17, int(impl.bitrev7(i)), 3329)
        if int(z[i]) != expect:
            out.append(f"ERROR :: Zeta[{i}]={int(z[i])} : zeta_d = {expect}\n")
        else:
            out.append(f"SUCCESS :: Zeta[{i}] = zeta_d\n")
    return "".join(out)


def pke_encrypt_decrypt_test(impl):  # PKE_EncryptDecrypt_test.c:14-47 (ML-KEM-512, d = r = 0..31, m[i] = i % 5)
    d = np.arange(32, dtype=np.uint8)[None, :]
    m = (np.arange(32) % 5).astype(np.uint8)[None, :]
    ek, dk = impl.pke_keygen(512, d)
    c = impl.pke_encrypt(512, ek, m, d)
    m2 = impl.pke_decrypt(512, dk, c)
    s = "Plaintext: " + "".join(f"{int(x)} " for x in m[0]) + "\n\n"
    s += "Ciphertext: " + "".join(f"{int(x)} " for x in c[0]) + "\n\n"
    s += "Plaintext: " + "".join(f"{int(x)} " for x in m2[0]) + "\n\n"
    return s


DRIVERS = {
    "BitRev7_test01": bitrev7_test01,
    "BitsAndBytes_test02": bits_and_bytes_test02,
    "EncodeDecode_test03": encode_decode_test03,
    "CompressDecompress_test04": compress_decompress_test04,
    "SampleNTT_test06": sample_ntt_test06,
    "SampleCBD_test07": sample_cbd_test07,
    "NTT_test08": ntt_test08,
    "ZetaLogic_test": zeta_logic_test,
    "PKE_EncryptDecrypt_test": pke_encrypt_decrypt_test,
}
