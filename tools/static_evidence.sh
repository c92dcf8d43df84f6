#!/bin/bash
# Static evidence for profiles/: registers / spills / shared memory of every kernel (nvcc -Xptxas -v) and SASS opcode
# histograms of the hot kernels split by issue pipe (cuobjdump -sass).  Runs without a GPU.
#   tools/static_evidence.sh r02
set -e
tag=${1:-r02}
cd "$(dirname "$0")/.."
out=profiles/ptxas_v_${tag}.txt
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden \
    -Xptxas -v -c -o /tmp/mlkem_b200_ptxas.o crystals-kyber_b200/csrc/mlkem_b200.cu 2>&1 | c++filt | \
python3 -c '
import re, sys
txt = sys.stdin.read()
rows = []
for m in re.finditer(r"Compiling entry function .(.*?). for .sm_100a.\n.*?Function properties for .*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes smem)?", txt, re.S):
    rows.append((m.group(1), int(m.group(5)), int(m.group(2)), int(m.group(3)), int(m.group(4)), int(m.group(6) or 0)))
print("# nvcc 12.9 -O3 -gencode arch=compute_100a,code=sm_100a -Xptxas -v, crystals-kyber_b200/csrc/mlkem_b200.cu")
print("# registers  stack  spill_st  spill_ld  static_smem  kernel")
for name, regs, stack, sst, sld, smem in sorted(rows, key=lambda r: r[0]):
    short = re.sub(r"\(.*", "", name)
    print(f"{regs:9d} {stack:6d} {sst:9d} {sld:9d} {smem:12d}  {short}")
print(f"# {len(rows)} kernels, {sum(1 for r in rows if r[3] or r[4])} with spills")
' > $out
echo "wrote $out"
hist=profiles/sass_hist_${tag}.txt
{
echo "# cuobjdump -sass opcode histograms (static instruction counts), tools/sass_static.py"
for pat in 'k_sample_matvec<.*3, 2, 2, 10, 4>, 1>' 'k_sample_matvec<.*3, 2, 2, 10, 4>, 2>' 'k_sample_matvec_list<.*3, 2, 2, 10, 4>, 1>' \
           'k_matvec_table<.*3, 2, 2, 10, 4>, 1>' 'k_decrypt<.*3, 2, 2, 10, 4>' 'k_encrypt_v<.*3, 2, 2, 10, 4>, false' 'k_noise<2, true, 21>' 'k_noise<2, false, 21>' \
           'k_encaps_HG<.*3, 2, 2, 10, 4>' 'k_decaps_J_select<.*3, 2, 2, 10, 4>, 21' 'k_encaps_HG_warp<.*3, 2, 2, 10, 4>' 'k_decaps_J_select_warp<.*3, 2, 2, 10, 4>, 21' 'k_ntt_batch' 'k_intt_batch' 'k_mulntt_batch'; do
    python3 tools/sass_static.py crystals-kyber_b200/libmlkem_b200.so "$pat"
done
echo
echo "# loop bodies (tools/sass_loop.py): the Keccak loop holds TWO rounds = 244 LOP3 + 116 SHF, i.e. 122 LOP3 + 58 SHF = 180 alu-pipe"
echo "# instructions per round; in the fused kernel: the Keccak loop, the three-block sampling loop around it (both parser variants),"
echo "# and phase 2 (one row per iteration)"
python3 tools/sass_loop.py crystals-kyber_b200/libmlkem_b200.so 'k_decaps_G<.*3, 2, 2, 10, 4>'
python3 tools/sass_loop.py crystals-kyber_b200/libmlkem_b200.so 'k_sample_matvec<.*3, 2, 2, 10, 4>, 1>'
python3 tools/sass_loop.py crystals-kyber_b200/libmlkem_b200.so 'k_matvec_table<.*3, 2, 2, 10, 4>, 1>'
echo "# one sponge per warp (small batches): the loop holds TWO rounds = 36 SHFL + 20 LOP3 + 8 SHF + 8 SEL"
python3 tools/sass_loop.py crystals-kyber_b200/libmlkem_b200.so 'k_encaps_HG_warp<.*3, 2, 2, 10, 4>'
} > $hist
echo "wrote $hist"
