// coissue_bench.cu -- can a scheduler of sm_100a issue another warp's instructions in the cycles a LOP3 stream leaves free?
//
// The alu pipe accepts one warp instruction every two cycles per scheduler, so a Keccak warp (pure LOP3 / SHF) uses every
// other issue slot.  This benchmark puts exactly two warps on every scheduler (one 256-thread block per SM, forced by dynamic
// shared memory): warp A runs a LOP3 stream of 8 independent chains, warp B runs the stream under test until A is done.
// Reported per scheduler: instructions per cycle of A alone, and of A and B when they share the scheduler.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o build/coissue_bench coissue_bench.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                        \
    do {                                                                                             \
        cudaError_t e_ = (x);                                                                        \
        if (e_ != cudaSuccess) {                                                                     \
            fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                                 \
        }                                                                                            \
    } while (0)

enum B { B_NONE = 0, B_IMAD, B_IMAD_CONST, B_IMADHI, B_IADD3, B_LDS, B_BUTTERFLY, B_LOP3, B_FFMA, B_FFMA_CONST, B_FADD, B_IMADWIDE, B_STS, B_ISETP_SEL, B_COUNT };
static const char *kName[B_COUNT] = {"none", "imad(3 regs)", "imad(const operand)", "imad.hi", "iadd3", "lds.u16", "lazy butterfly (IMAD.HI, 2 IMAD, 2 IADD)", "lop3", "ffma(3 regs)", "ffma(const operand)", "fadd", "imad.wide", "sts.u16", "isetp+sel"};

__constant__ uint32_t c_k[4];

template <int BOP>
__global__ void __launch_bounds__(256, 1) k(uint32_t *out, int iters, unsigned long long *res) {
    extern __shared__ uint32_t smem[];
    volatile int *done = reinterpret_cast<volatile int *>(smem);
    const int warp = threadIdx.x >> 5;
    const bool is_a = warp < 4;  // warps 0..3 -> schedulers 0..3 (A), warps 4..7 -> schedulers 0..3 (B)
    if (threadIdx.x == 0) *done = 0;
    __syncthreads();
    uint32_t x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    uint32_t m = out[0] | 0x10001u, s = (out[1] & 7) + 3, q = out[2] + 3329u;
    unsigned long long t0 = clock64();
    unsigned long long count = 0;
    if (is_a) {
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int r = 0; r < 4; r++) {
#define L3(a) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(m), "r"(s));
                L3(x0) L3(x1) L3(x2) L3(x3) L3(x4) L3(x5) L3(x6) L3(x7)
            }
        }
        count = (unsigned long long)iters * 32;
        __syncwarp();
        if ((threadIdx.x & 31) == 0) atomicAdd((int *)done, 1);
    } else if (BOP != B_NONE) {
        const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem + 64 + threadIdx.x);
        while (*done < 4) {
#pragma unroll
            for (int r = 0; r < 4; r++) {
                if (BOP == B_IMAD) {
#define IM(a) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(m), "r"(s));
                    IM(x0) IM(x1) IM(x2) IM(x3) IM(x4) IM(x5) IM(x6) IM(x7)
                } else if (BOP == B_IMAD_CONST) {
#define IC(a) a = a * c_k[1] + 12345u;
                    IC(x0) IC(x1) IC(x2) IC(x3) IC(x4) IC(x5) IC(x6) IC(x7)
                } else if (BOP == B_IMADHI) {
#define IH(a) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(m), "r"(s));
                    IH(x0) IH(x1) IH(x2) IH(x3) IH(x4) IH(x5) IH(x6) IH(x7)
                } else if (BOP == B_IADD3) {
#define IA(a) asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(m));
                    IA(x0) IA(x1) IA(x2) IA(x3) IA(x4) IA(x5) IA(x6) IA(x7)
                } else if (BOP == B_LDS) {
#define LS(a) { uint16_t t_; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(t_) : "r"(sa + ((a) & 0)) : "memory"); a += t_; }
                    LS(x0) LS(x1) LS(x2) LS(x3) LS(x4) LS(x5) LS(x6) LS(x7)
                } else if (BOP == B_FFMA) {
#define FF(a) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(a) : "r"(m), "r"(s));
                    FF(x0) FF(x1) FF(x2) FF(x3) FF(x4) FF(x5) FF(x6) FF(x7)
                } else if (BOP == B_FFMA_CONST) {
#define FC(a) a = __float_as_uint(fmaf(__uint_as_float(a), __uint_as_float(c_k[2]), 1.5f));
                    FC(x0) FC(x1) FC(x2) FC(x3) FC(x4) FC(x5) FC(x6) FC(x7)
                } else if (BOP == B_FADD) {
#define FA(a) asm volatile("add.rn.f32 %0, %0, %1;" : "+r"(a) : "r"(m));
                    FA(x0) FA(x1) FA(x2) FA(x3) FA(x4) FA(x5) FA(x6) FA(x7)
                } else if (BOP == B_IMADWIDE) {
#define IW(a) { uint32_t hi_; asm volatile("{.reg .b64 t; mul.wide.u32 t, %0, %2; mov.b64 {%0, %1}, t;}" : "+r"(a), "=r"(hi_) : "r"(m)); a ^= hi_ & 0; }
                    IW(x0) IW(x1) IW(x2) IW(x3) IW(x4) IW(x5) IW(x6) IW(x7)
                } else if (BOP == B_STS) {
#define SS(a) asm volatile("st.shared.u16 [%0], %1;" ::"r"(sa), "h"((uint16_t)(a)) : "memory");
                    SS(x0) SS(x1) SS(x2) SS(x3) SS(x4) SS(x5) SS(x6) SS(x7)
                } else if (BOP == B_ISETP_SEL) {
#define IS(a) asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; @p add.u32 %0, %0, %2;}" : "+r"(a) : "r"(m), "r"(s));
                    IS(x0) IS(x1) IS(x2) IS(x3) IS(x4) IS(x5) IS(x6) IS(x7)
                } else if (BOP == B_LOP3) {
                    L3(x0) L3(x1) L3(x2) L3(x3) L3(x4) L3(x5) L3(x6) L3(x7)
                } else if (BOP == B_BUTTERFLY) {
                    // four lazy Cooley-Tukey butterflies on (x0,x4) (x1,x5) (x2,x6) (x3,x7): 5 instructions each ... x2 rounds
#define BF(a, b)                                                                                          \
    {                                                                                                     \
        uint32_t qh_, t_;                                                                                 \
        asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(qh_) : "r"(b), "r"(m));                              \
        asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(t_) : "r"(b), "r"(s));                               \
        asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(t_) : "r"(qh_), "r"(q));                         \
        asm volatile("sub.u32 %0, %1, %2;" : "=r"(b) : "r"(a), "r"(t_));                                  \
        asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(t_));                                          \
    }
                    BF(x0, x4) BF(x1, x5) BF(x2, x6) BF(x3, x7)
                    BF(x0, x2) BF(x1, x3) BF(x4, x6) BF(x5, x7)
                }
            }
            count += (BOP == B_BUTTERFLY) ? 4 * 40 : (BOP == B_ISETP_SEL) ? 64 : 32;
        }
    }
    unsigned long long t1 = clock64();
    uint32_t acc = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
    if (acc == 0x12345678u) out[3] = acc;
    if ((threadIdx.x & 31) == 0) {
        res[(blockIdx.x * 8 + warp) * 2] = t1 - t0;
        res[(blockIdx.x * 8 + warp) * 2 + 1] = count;
    }
}

template <int BOP>
static void run(int sms, uint32_t *d_out, unsigned long long *d_res) {
    const int iters = 20000;
    const size_t smem = 200 * 1024;
    CK(cudaFuncSetAttribute(k<BOP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<BOP><<<sms, 256, smem>>>(d_out, 100, d_res);
    CK(cudaDeviceSynchronize());
    k<BOP><<<sms, 256, smem>>>(d_out, iters, d_res);
    CK(cudaDeviceSynchronize());
    std::vector<unsigned long long> h((size_t)sms * 16);
    CK(cudaMemcpy(h.data(), d_res, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    double a_ipc = 0, b_ipc = 0;
    for (int b = 0; b < sms; b++)
        for (int w = 0; w < 8; w++) {
            double cyc = (double)h[(b * 8 + w) * 2], cnt = (double)h[(b * 8 + w) * 2 + 1];
            (w < 4 ? a_ipc : b_ipc) += cnt / cyc;
        }
    a_ipc /= sms * 4;
    b_ipc /= sms * 4;
    printf("{\"b_stream\": \"%s\", \"lop3_warp_ipc\": %.4f, \"b_warp_ipc\": %.4f, \"scheduler_ipc\": %.4f}\n", kName[BOP], a_ipc, b_ipc, a_ipc + b_ipc);
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    uint32_t *d_out;
    unsigned long long *d_res;
    CK(cudaMalloc(&d_out, 64));
    CK(cudaMemset(d_out, 0, 64));
    CK(cudaMalloc(&d_res, sizeof(unsigned long long) * prop.multiProcessorCount * 16));
    uint32_t hk[4] = {1, 1 << 12, 3, 5};
    CK(cudaMemcpyToSymbol(c_k, hk, sizeof hk));
    run<B_NONE>(prop.multiProcessorCount, d_out, d_res);
    run<B_LOP3>(prop.multiProcessorCount, d_out, d_res);
    run<B_IMAD>(prop.multiProcessorCount, d_out, d_res);
    run<B_IMAD_CONST>(prop.multiProcessorCount, d_out, d_res);
    run<B_IMADHI>(prop.multiProcessorCount, d_out, d_res);
    run<B_IADD3>(prop.multiProcessorCount, d_out, d_res);
    run<B_LDS>(prop.multiProcessorCount, d_out, d_res);
    run<B_BUTTERFLY>(prop.multiProcessorCount, d_out, d_res);
    run<B_FFMA>(prop.multiProcessorCount, d_out, d_res);
    run<B_FFMA_CONST>(prop.multiProcessorCount, d_out, d_res);
    run<B_FADD>(prop.multiProcessorCount, d_out, d_res);
    run<B_IMADWIDE>(prop.multiProcessorCount, d_out, d_res);
    run<B_STS>(prop.multiProcessorCount, d_out, d_res);
    run<B_ISETP_SEL>(prop.multiProcessorCount, d_out, d_res);
    return 0;
}
