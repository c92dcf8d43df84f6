"""The reference's SHA-3 test flow (Test_Archive/SHA/sha_testing.sh) restated for a given test05 binary.

sha_testing.sh parses every file of Test_Examples/SHA with sha_ex_psr.pl into bin_msg.txt (the message as ASCII 0/1
digits) and exp_output.txt (the expected hex string), runs `./test05 Hash <d>` or `./test05 XOF <strength> 4096`
(Test_Archive/SHA/SHA_Testing.c: reads bin_msg.txt, calls sha3_b + b2h, writes act_output.txt) and compares the two
case-insensitively.  Here the parsed examples come from tests/golden/sha_examples.json (made from the same 16 files by
tests/golden/make_golden.py), so the flow runs where /root/reference does not exist; the binary is the UNMODIFIED
SHA_Testing.c, linked either with the reference's sha3.c (CPU test) or with libmlkem_b200.so (GPU test).
"""
import os
import subprocess


def run(exe, examples, workdir):
    """Returns the list of example names that passed; raises AssertionError on the first mismatch."""
    passed = []
    for ex in examples:
        fnct, dlen = ex["name"].split("_")[0].split("-")  # "XOF-128_Msg1605" -> XOF, 128  (sha_testing.sh:13-17)
        with open(os.path.join(workdir, "bin_msg.txt"), "w") as f:
            f.write(ex["bits"])
        args = [exe, fnct, dlen] + (["4096"] if fnct == "XOF" else [])  # sha_testing.sh:19-24
        out = subprocess.run(args, cwd=workdir, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
        assert out.returncode == 0, (ex["name"], out.stdout[-200:], out.stderr[-200:])
        with open(os.path.join(workdir, "act_output.txt")) as f:
            act = f.read()
        assert act.lower() == ex["hex"].lower(), (ex["name"], act[:32], ex["hex"][:32])  # sha_testing.sh:28
        for name in ("bin_msg.txt", "act_output.txt"):
            os.remove(os.path.join(workdir, name))
        passed.append(ex["name"])
    return passed
