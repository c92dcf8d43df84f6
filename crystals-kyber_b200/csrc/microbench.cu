// microbench.cu -- measures the INT32 roofline denominators of one B200 (sm_100a).
//
// MEASURED_PEAKS.json (driver-written) only has HBM GB/s and dense bf16 TFLOP/s.  The ML-KEM hot path is
// bound by the 32-bit integer pipes, so the denominator for `roofline.frac` has to be
// measured here: issue rates, in thread-operations per clock per SM, of
//   LOP3 / SHF   (alu pipe: what Keccak-f[1600] is made of)
//   IMAD / IMAD.HI / IMAD.WIDE (fma pipe: what Barrett/Shoup modular multiplication is made of)
//   LOP3 + IMAD interleaved (both pipes)
// Each test is a fully occupied grid of threads running 8 independent dependency chains of the instruction
// under test; elapsed SM cycles come from clock64() so the per-clock figure does not depend on DVFS, and
// the wall-clock figure (CUDA events) gives the sustained op/s at whatever clock the part held.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o build/microbench microbench.cu
// Output: one JSON object on stdout.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                   \
        }                                                                              \
    } while (0)

enum Op { OP_LOP3 = 0, OP_SHF, OP_IMAD, OP_IMADHI, OP_IMADWIDE, OP_MIX_LOP3_IMAD, OP_MIX_SHF_LOP3, OP_IADD3, OP_COUNT };
static const char *kOpName[OP_COUNT] = {"lop3", "shf", "imad", "imad_hi", "imad_wide", "lop3+imad", "shf+lop3", "iadd3"};
// thread-operations per loop iteration (8 chains x 4 repeats, mixes count both instructions)
static const int kOpsPerIter[OP_COUNT] = {32, 32, 32, 32, 32, 64, 64, 32};

template <int OP>
__global__ void __launch_bounds__(1024, 2) rate_kernel(uint32_t *out, int iters, unsigned long long *cycles) {
    uint32_t x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    uint32_t y0 = blockIdx.x | 1, y1 = y0 + 2, y2 = y0 + 4, y3 = y0 + 6, y4 = y0 + 8, y5 = y0 + 10, y6 = y0 + 12, y7 = y0 + 14;
    uint32_t m = out[0] | 0x10001u, s = (out[1] & 7) + 3;  // run-time values: nothing can be strength-reduced
    unsigned long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            if (OP == OP_LOP3 || OP == OP_MIX_LOP3_IMAD || OP == OP_MIX_SHF_LOP3) {
#define L3(a) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(m), "r"(s));
                L3(x0) L3(x1) L3(x2) L3(x3) L3(x4) L3(x5) L3(x6) L3(x7)
            }
            if (OP == OP_SHF || OP == OP_MIX_SHF_LOP3) {
#define SF(a) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(m), "r"(s));
                if (OP == OP_SHF) { SF(x0) SF(x1) SF(x2) SF(x3) SF(x4) SF(x5) SF(x6) SF(x7) }
                else { SF(y0) SF(y1) SF(y2) SF(y3) SF(y4) SF(y5) SF(y6) SF(y7) }
            }
            if (OP == OP_IMAD) {
#define IM(a) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(m), "r"(s));
                IM(x0) IM(x1) IM(x2) IM(x3) IM(x4) IM(x5) IM(x6) IM(x7)
            }
            if (OP == OP_MIX_LOP3_IMAD) { IM(y0) IM(y1) IM(y2) IM(y3) IM(y4) IM(y5) IM(y6) IM(y7) }
            if (OP == OP_IMADHI) {
#define IH(a) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(m), "r"(s));
                IH(x0) IH(x1) IH(x2) IH(x3) IH(x4) IH(x5) IH(x6) IH(x7)
            }
            if (OP == OP_IMADWIDE) {
#define IW(a, b) asm volatile("{.reg .b64 t; mov.b64 t, {%0, %1}; mad.wide.u32 t, %0, %2, t; mov.b64 {%0, %1}, t;}" : "+r"(a), "+r"(b) : "r"(m));
                IW(x0, y0) IW(x1, y1) IW(x2, y2) IW(x3, y3) IW(x4, y4) IW(x5, y5) IW(x6, y6) IW(x7, y7)
            }
            if (OP == OP_IADD3) {
#define IA(a) asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(m));
                IA(x0) IA(x1) IA(x2) IA(x3) IA(x4) IA(x5) IA(x6) IA(x7)
            }
        }
    }
    unsigned long long t1 = clock64();
    uint32_t acc = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7 ^ y0 ^ y1 ^ y2 ^ y3 ^ y4 ^ y5 ^ y6 ^ y7;
    if (acc == 0x12345678u) out[2] = acc;  // keep the chains alive
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
static void run_rate(int sms, uint32_t *d_out, unsigned long long *d_cyc, bool last) {
    const int iters = 4096, blocks = sms * 2, threads = 1024;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    rate_kernel<OP><<<blocks, threads>>>(d_out, 64, d_cyc);  // warm-up
    CK(cudaDeviceSynchronize());
    float best_ms = 1e30f;
    double cyc_mean = 0;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(e0));
        rate_kernel<OP><<<blocks, threads>>>(d_out, iters, d_cyc);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best_ms) {
            best_ms = ms;
            std::vector<unsigned long long> h(blocks);
            CK(cudaMemcpy(h.data(), d_cyc, blocks * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            cyc_mean = 0;
            for (auto c : h) cyc_mean += (double)c;
            cyc_mean /= blocks;
        }
    }
    // two resident 1024-thread CTAs per SM run concurrently for cyc_mean cycles
    double ops_per_sm = 2.0 * threads * (double)iters * kOpsPerIter[OP];
    double per_clk_sm = ops_per_sm / cyc_mean;
    double total_ops = (double)blocks * threads * (double)iters * kOpsPerIter[OP];
    printf("  \"%s\": {\"thread_ops_per_clk_per_sm\": %.2f, \"tera_ops_per_s\": %.3f, \"ms\": %.4f}%s\n", kOpName[OP],
           per_clk_sm, total_ops / (best_ms * 1e-3) / 1e12, best_ms, last ? "" : ",");
    CK(cudaEventDestroy(e0));
    CK(cudaEventDestroy(e1));
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    uint32_t *d_out;
    unsigned long long *d_cyc;
    CK(cudaMalloc(&d_out, 64));
    CK(cudaMemset(d_out, 0, 64));
    CK(cudaMalloc(&d_cyc, sizeof(unsigned long long) * prop.multiProcessorCount * 2));
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    printf("{\n  \"gpu\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\", \"max_sm_khz\": %d,\n", prop.name, prop.multiProcessorCount,
           prop.major, prop.minor, clk_khz);
    run_rate<OP_LOP3>(prop.multiProcessorCount, d_out, d_cyc, false);
    run_rate<OP_SHF>(prop.multiProcessorCount, d_out, d_cyc, false);
    run_rate<OP_IADD3>(prop.multiProcessorCount, d_out, d_cyc, false);
    run_rate<OP_IMAD>(prop.multiProcessorCount, d_out, d_cyc, false);
    run_rate<OP_IMADHI>(prop.multiProcessorCount, d_out, d_cyc, false);
    run_rate<OP_IMADWIDE>(prop.multiProcessorCount, d_out, d_cyc, false);
    run_rate<OP_MIX_SHF_LOP3>(prop.multiProcessorCount, d_out, d_cyc, false);
    run_rate<OP_MIX_LOP3_IMAD>(prop.multiProcessorCount, d_out, d_cyc, true);
    printf("}\n");
    return 0;
}
