#!/usr/bin/env python3
"""Generate tests/golden/*.json from the UNMODIFIED reference at /root/reference.

Run in the build container (needs /root/reference and gcc):

    python tests/golden/make_golden.py

Two kinds of fixtures are written (data only -- no reference code enters the repo):

1. archive_stdout.json -- sha256 of the stdout of every deterministic Test_Archive driver of the
   reference, built at HEAD with the recipe of SURVEY.md Appendix B (the driver is appended to a
   translation unit that #includes ml_kem.c, and linked with sha3.c compiled with the makefile's flags).
2. ref_vectors.json -- inputs and outputs of the reference's own functions (through oracle/ref_shim.c)
   on seeded inputs: every hot-path function, all three parameter sets, incl. the implicit-rejection path.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import Reference, build, sizes  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
GCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"


def archive_stdout():
    drivers = ["BitRev7_test01", "BitsAndBytes_test02", "EncodeDecode_test03", "CompressDecompress_test04",
               "SampleNTT_test06", "SampleCBD_test07", "NTT_test08", "ZetaLogic_test", "PKE_EncryptDecrypt_test"]
    res = {}
    with tempfile.TemporaryDirectory() as td:
        subprocess.run([GCC, "-Wall", "-g", "-w", "-c", f"{REF}/sha3.c", "-o", f"{td}/sha3.o"], check=True)
        os.makedirs(f"{td}/inc")
        os.symlink(f"{REF}/ml_kem.h", f"{td}/inc/ml-kem.h")  # test01 includes the header's old name
        for drv in drivers:
            src = f"{REF}/Test_Archive/{drv}.c"
            if drv == "BitsAndBytes_test02":  # uses the obsolete member name `.o` (now `.b`)
                txt = open(src).read().replace(".o ", ".b ").replace(".o)", ".b)").replace(".o;", ".b;").replace(".o,", ".b,")
                src = f"{td}/{drv}.c"
                open(src, "w").write(txt)
            open(f"{td}/w.c", "w").write(f'#include "{REF}/ml_kem.c"\n#include "{src}"\n')
            exe = f"{td}/{drv}"
            subprocess.run([GCC, "-w", "-g", f"-I{REF}", f"-I{td}/inc", f"{td}/w.c", f"{td}/sha3.o", "-o", exe], check=True)
            out = subprocess.run([exe], stdout=subprocess.PIPE).stdout
            res[drv] = {"sha256": hashlib.sha256(out).hexdigest(), "bytes": len(out)}
            if len(out) < 4096:
                res[drv]["stdout"] = out.decode()
    # the corrected ML-KEM-768 round-trip driver of BASELINE config 1 (public API, random seeds):
    return res


def ref_vectors(r: Reference):
    rng = np.random.default_rng(20261018)
    v = {"sizes": r.sizes()}
    v["init"] = {str(ps): list(r.init(ps)[1]) for ps in (512, 768, 1024)}
    v["init_bad"] = r.init(999)[0]
    v["bitrev7"] = [r.bitrev7(i) for i in range(128)]
    # Compress / Decompress: exhaustive over all 12-bit inputs (hash) for d = 1..12
    comp, decomp = {}, {}
    for d in range(1, 13):
        comp[str(d)] = hashlib.sha256(np.array([r.compress(x, d) for x in range(4096)], np.uint16).tobytes()).hexdigest()
        decomp[str(d)] = hashlib.sha256(np.array([r.decompress(x, d) for x in range(1 << d)], np.uint16).tobytes()).hexdigest()
    v["compress_sha256"], v["decompress_sha256"] = comp, decomp
    # ByteEncode / ByteDecode
    enc = []
    for d in (1, 4, 5, 10, 11, 12):
        F = rng.integers(0, 1 << d, 256, dtype=np.uint16)
        B = r.byte_encode(F, d)
        assert (r.byte_decode(B, d) == F).all()
        enc.append({"d": d, "F": F.tobytes().hex(), "B": B.tobytes().hex()})
    raw = rng.integers(0, 256, 384, dtype=np.uint8)  # arbitrary bytes incl. values >= q (D4)
    enc.append({"d": 12, "decode_only": True, "B": raw.tobytes().hex(), "F": r.byte_decode(raw, 12).tobytes().hex()})
    raw = np.full(384, 0xFF, np.uint8)
    enc.append({"d": 12, "decode_only": True, "B": raw.tobytes().hex(), "F": r.byte_decode(raw, 12).tobytes().hex()})
    v["encode"] = enc
    # SampleNTT: the 7 seeds of SampleNTT_test06.c plus random seeds
    sn = []
    seeds = [bytes(((it * i + i) & 0xFF) for i in range(34)) for it in range(7)]
    seeds += [rng.integers(0, 256, 34, dtype=np.uint8).tobytes() for _ in range(9)]
    for s in seeds:
        a, after = r.sample_ntt(s)
        sn.append({"seed": s.hex(), "seed_after": after.hex(), "a": a.tobytes().hex()})
    v["sample_ntt"] = sn
    # SamplePolyCBD
    cb = []
    for eta in (2, 3):
        for t in range(3):
            data = bytes(range(64 * eta)) if t == 0 else rng.integers(0, 256, 64 * eta, dtype=np.uint8).tobytes()
            cb.append({"eta": eta, "B": data.hex(), "f": r.sample_cbd(data, eta).tobytes().hex()})
    v["cbd"] = cb
    # hashes on ML-KEM-sized inputs
    hs = []
    for name, fn, lens in (("H", r.H, (0, 32, 135, 136, 137, 800, 1184, 1568)), ("G", r.G, (33, 64, 71, 72, 73)),
                           ("J", r.J, (32, 167, 168, 169, 800, 1120, 1600))):
        for n in lens:
            data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
            hs.append({"fn": name, "in": data.hex(), "out": fn(data).hex()})
    for eta in (2, 3):
        for b in (0, 1, 7, 255):
            s = rng.integers(0, 256, 32, dtype=np.uint8).tobytes()
            hs.append({"fn": "PRF", "s": s.hex(), "b": b, "eta": eta, "out": r.PRF(s, b, eta).hex()})
    v["hash"] = hs
    # ring arithmetic
    ring = []
    for t in range(6):
        f = rng.integers(0, 3329, 256, dtype=np.uint16)
        g = rng.integers(0, 3329, 256, dtype=np.uint16)
        if t >= 4:  # non-canonical 12-bit values in the first operand (what ByteDecode12 can hand over, D4)
            f = rng.integers(0, 4096, 256, dtype=np.uint16)
        ring.append({"f": f.tobytes().hex(), "g": g.tobytes().hex(),
                     "ntt_f": r.ntt(f).tobytes().hex() if t < 4 else None,
                     "intt_f": r.intt(f).tobytes().hex() if t < 4 else None,
                     "mul": r.multiply_ntts(f, g).tobytes().hex(),
                     "add": r.poly_add(f % 3329, g).tobytes().hex(),
                     "sub": r.poly_sub(f % 3329, g).tobytes().hex()})
    v["ring"] = ring
    v["basecase"] = [{"in": [a0, a1, b0, b1, gm], "out": list(r.basecase_multiply(a0, a1, b0, b1, gm))}
                     for a0, a1, b0, b1, gm in ([1, 2, 3, 4, 17], [3328, 3328, 3328, 3328, 3312], [4095, 4095, 3328, 3328, 2761],
                                                [0, 0, 0, 0, 17], [1234, 4000, 77, 3000, 568])]
    # K-PKE + ML-KEM internal, all parameter sets
    kem = []
    for ps in (512, 768, 1024):
        for t in range(3):
            if t == 0:
                d, z, m = bytes(range(32)), bytes(range(32, 64)), bytes(range(64, 96))
            else:
                d, z, m = (rng.integers(0, 256, 32, dtype=np.uint8).tobytes() for _ in range(3))
            ek, dk = r.keygen_internal(ps, d, z)
            pek, pdk = r.pke_keygen(ps, d)
            assert pek == ek and pdk == dk[: len(pdk)]
            c, K = r.encaps_internal(ps, ek, m)
            Kd = r.decaps_internal(ps, dk, c)
            assert Kd == K
            cb = bytearray(c)
            cb[5] ^= 1
            Krej = r.decaps_internal(ps, dk, bytes(cb))
            rr = rng.integers(0, 256, 32, dtype=np.uint8).tobytes()
            cp = r.pke_encrypt(ps, ek, m, rr)
            mp = r.pke_decrypt(ps, pdk, cp)
            assert mp == m
            rec = {"set": ps, "d": d.hex(), "z": z.hex(), "m": m.hex(), "K": K.hex(), "K_rej": Krej.hex(),
                   "ek_sha256": hashlib.sha256(ek).hexdigest(), "dk_sha256": hashlib.sha256(dk).hexdigest(),
                   "c_sha256": hashlib.sha256(c).hexdigest(), "pke_r": rr.hex(),
                   "pke_c_sha256": hashlib.sha256(cp).hexdigest()}
            if t == 0:
                rec.update({"ek": ek.hex(), "dk": dk.hex(), "c": c.hex()})
            kem.append(rec)
    v["kem"] = kem
    # all-0xFF encapsulation key (D4: the modulus check of KEM_Encaps cannot fail)
    ek_ff = bytes([0xFF]) * sizes(768)["ek"]
    m = bytes(range(32))
    c, K = r.encaps_internal(768, ek_ff, m)
    v["ek_all_ff_768"] = {"m": m.hex(), "c_sha256": hashlib.sha256(c).hexdigest(), "K": K.hex()}
    # the PKE KAT of PKE_EncryptDecrypt_test.c (ML-KEM-512, d = r = 0..31, m[i] = i % 5)
    d = bytes(range(32))
    m5 = bytes(i % 5 for i in range(32))
    ek, dk = r.pke_keygen(512, d)
    c = r.pke_encrypt(512, ek, m5, d)
    assert r.pke_decrypt(512, dk, c) == m5
    v["pke_test10"] = {"c": c.hex(), "rho": ek[-32:].hex()}
    return v


def sha_examples(r: Reference):
    """The 16 NIST example files of Test_Examples/SHA, parsed like Test_Archive/SHA/sha_ex_psr.pl does (message bits
    after "Msg as bit string", expected value after "Hash val is" / "Output val is"), each checked here against the
    compiled reference the way sha_testing.sh does (XOFs squeeze 4096 bits), plus reference outputs for message
    lengths around the block boundary, including the lengths where the reference's padding deviates from FIPS 202."""
    import glob
    import re

    out = {"examples": [], "boundary": []}
    for path in sorted(glob.glob(f"{REF}/Test_Examples/SHA/*.txt")):
        name = os.path.basename(path)[:-4]
        txt = open(path).read()
        m = re.search(r"Msg as bit string\n(.*?)\n\s*\n", txt, re.S)
        bits = re.sub(r"\s", "", m.group(1)) if m else ""
        bits = bits if re.fullmatch(r"[01]+", bits or "x") else ""
        m = re.search(r"(?:Hash val is|Output val is)\n(.*)$", txt, re.S)
        expect = re.sub(r"\s", "", m.group(1)).lower()
        kind, dlen = name.split("_")[0].split("-")
        dlen = int(dlen)
        if kind == "XOF":
            sfx, c, d = [1, 1, 1, 1], 2 * dlen, 4096
        else:
            sfx, c, d = [0, 1, 0, 0], 2 * dlen, dlen
        got = r.sha3_bits([int(b) for b in bits], sfx, c, d)
        hexs = np.packbits(got, bitorder="little").tobytes().hex()
        assert hexs == expect, name
        out["examples"].append({"name": name, "bits": bits, "sfx": sfx, "c": c, "d": d, "hex": expect})
    rng = np.random.default_rng(202)
    for sfx, c, d in (([0, 1, 0, 0], 448, 224), ([0, 1, 0, 0], 512, 256), ([0, 1, 0, 0], 768, 384), ([0, 1, 0, 0], 1024, 512),
                      ([1, 1, 1, 1], 256, 1400), ([1, 1, 1, 1], 512, 1100)):
        rr, slen = 1600 - c, (4 if sfx[2] else 2)
        for n in (rr - slen - 3, rr - slen - 2, rr - slen - 1, rr - slen, rr - slen + 1, 2 * rr - slen - 2, 2 * rr - slen - 1, 13, 777):
            bits = rng.integers(0, 2, n, dtype=np.uint8)
            got = r.sha3_bits(bits, sfx, c, d)
            out["boundary"].append({"bits": "".join(map(str, bits)), "sfx": sfx, "c": c, "d": d,
                                    "out_bits_sha256": hashlib.sha256(got.tobytes()).hexdigest(),
                                    "quirk": (n + slen + 2) % rr == 0})
    out["sha3_s"] = [{"text": t, "hex": r.sha3_s(t.encode(), [0, 1, 0, 0], 512, 256).hex()} for t in ("", "abc", "CRYSTALS-Kyber on B200")]
    return out


def main():
    build()
    r = Reference()
    arch = archive_stdout()
    json.dump(arch, open(os.path.join(OUT, "archive_stdout.json"), "w"), indent=1, sort_keys=True)
    vec = ref_vectors(r)
    json.dump(vec, open(os.path.join(OUT, "ref_vectors.json"), "w"), indent=0, sort_keys=True)
    json.dump(sha_examples(r), open(os.path.join(OUT, "sha_examples.json"), "w"), indent=0, sort_keys=True)
    for k, val in arch.items():
        print(f"{k:28s} {val['sha256'][:16]} {val['bytes']} B")
    print("wrote", OUT)


if __name__ == "__main__":
    main()
