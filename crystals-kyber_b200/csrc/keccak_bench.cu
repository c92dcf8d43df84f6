#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__constant__ uint2 RC[24] = {
{0x00000001u,0x00000000u},{0x00008082u,0x00000000u},{0x0000808au,0x80000000u},{0x80008000u,0x80000000u},{0x0000808bu,0x00000000u},{0x80000001u,0x00000000u},
{0x80008081u,0x80000000u},{0x00008009u,0x80000000u},{0x0000008au,0x00000000u},{0x00000088u,0x00000000u},{0x80008009u,0x00000000u},{0x8000000au,0x00000000u},
{0x8000808bu,0x00000000u},{0x0000008bu,0x80000000u},{0x00008089u,0x80000000u},{0x00008003u,0x80000000u},{0x00008002u,0x80000000u},{0x00000080u,0x80000000u},
{0x0000800au,0x00000000u},{0x8000000au,0x80000000u},{0x80008081u,0x80000000u},{0x00008080u,0x80000000u},{0x80000001u,0x00000000u},{0x80008008u,0x80000000u}};
__constant__ uint32_t POW2[32];
struct L { uint32_t lo, hi; };
template<int N, int MODE> __device__ __forceinline__ L rotl(L x){
  L r;
  if (N==0) return x;
  if (N==32){ r.lo=x.hi; r.hi=x.lo; return r; }
  if (MODE==0) {
    if (N<32){ r.hi=__funnelshift_l(x.lo,x.hi,N); r.lo=__funnelshift_l(x.hi,x.lo,N); return r; }
    r.hi=__funnelshift_l(x.hi,x.lo,N-32); r.lo=__funnelshift_l(x.lo,x.hi,N-32); return r;
  } else {
    // rotate through the fma pipe: two IMAD.WIDE with a run-time 2^n multiplier
    constexpr int n = N & 31;
    uint32_t lo = N<32 ? x.lo : x.hi, hi = N<32 ? x.hi : x.lo;
    uint32_t p = POW2[n];
    unsigned long long t = (unsigned long long)lo * p;               // (lo<<n) | (lo>>(32-n))<<32
    unsigned long long sw = (t >> 32) | (t << 32);                    // swap halves
    unsigned long long u = (unsigned long long)hi * p + sw;          // lo: hi<<n + lo>>(32-n) ; hi: hi>>(32-n) + lo<<n
    r.hi = (uint32_t)u; r.lo = (uint32_t)(u >> 32);
    return r;
  }
}
__device__ __forceinline__ L x3(L a,L b,L c){ L r; r.lo=a.lo^b.lo^c.lo; r.hi=a.hi^b.hi^c.hi; return r; }
__device__ __forceinline__ L x5(L a,L b,L c,L d,L e){ L r; r.lo=a.lo^b.lo^c.lo^d.lo^e.lo; r.hi=a.hi^b.hi^c.hi^d.hi^e.hi; return r; }
__device__ __forceinline__ L chi(L a,L b,L c){ L r; r.lo=a.lo^(~b.lo&c.lo); r.hi=a.hi^(~b.hi&c.hi); return r; }
// MODE bit0: theta rot via imad; MODE bits: number of rho lanes rotated via imad (0..24)
template<int UU, int NI> __device__ __forceinline__ void keccak(L a[25]) {
  #pragma unroll UU
  for (int r=0;r<24;r++){
    L c0=x5(a[0],a[5],a[10],a[15],a[20]), c1=x5(a[1],a[6],a[11],a[16],a[21]), c2=x5(a[2],a[7],a[12],a[17],a[22]), c3=x5(a[3],a[8],a[13],a[18],a[23]), c4=x5(a[4],a[9],a[14],a[19],a[24]);
    L r0=rotl<1,0>(c0),r1=rotl<1,0>(c1),r2=rotl<1,0>(c2),r3=rotl<1,0>(c3),r4=rotl<1,0>(c4);
    #define TH(x,cm,rp) a[x]=x3(a[x],cm,rp); a[x+5]=x3(a[x+5],cm,rp); a[x+10]=x3(a[x+10],cm,rp); a[x+15]=x3(a[x+15],cm,rp); a[x+20]=x3(a[x+20],cm,rp);
    TH(0,c4,r1) TH(1,c0,r2) TH(2,c1,r3) TH(3,c2,r4) TH(4,c3,r0)
    L b[25];
    #define M(i) ((i) < NI ? 1 : 0)
    b[0]=a[0]; b[10]=rotl<1,M(0)>(a[1]); b[20]=rotl<62,M(1)>(a[2]); b[5]=rotl<28,M(2)>(a[3]); b[15]=rotl<27,M(3)>(a[4]);
    b[16]=rotl<36,M(4)>(a[5]); b[1]=rotl<44,M(5)>(a[6]); b[11]=rotl<6,M(6)>(a[7]); b[21]=rotl<55,M(7)>(a[8]); b[6]=rotl<20,M(8)>(a[9]);
    b[7]=rotl<3,M(9)>(a[10]); b[17]=rotl<10,M(10)>(a[11]); b[2]=rotl<43,M(11)>(a[12]); b[12]=rotl<25,M(12)>(a[13]); b[22]=rotl<39,M(13)>(a[14]);
    b[23]=rotl<41,M(14)>(a[15]); b[8]=rotl<45,M(15)>(a[16]); b[18]=rotl<15,M(16)>(a[17]); b[3]=rotl<21,M(17)>(a[18]); b[13]=rotl<8,M(18)>(a[19]);
    b[14]=rotl<18,M(19)>(a[20]); b[24]=rotl<2,M(20)>(a[21]); b[9]=rotl<61,M(21)>(a[22]); b[19]=rotl<56,M(22)>(a[23]); b[4]=rotl<14,M(23)>(a[24]);
    #pragma unroll
    for(int y=0;y<25;y+=5){
      #pragma unroll
      for(int x=0;x<5;x++) a[y+x]=chi(b[y+x],b[y+(x+1)%5],b[y+(x+2)%5]);
    }
    a[0].lo^=RC[r].x; a[0].hi^=RC[r].y;
  }
}
template<int UU, int NI, int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB) kk(uint2* io, int iters){
  L a[25];
  size_t n=(size_t)gridDim.x*blockDim.x; size_t t=(size_t)blockIdx.x*blockDim.x+threadIdx.x;
  #pragma unroll
  for(int i=0;i<25;i++){ uint2 v=io[i*n+t]; a[i].lo=v.x; a[i].hi=v.y; }
  for(int it=0;it<iters;it++) keccak<UU,NI>(a);
  #pragma unroll
  for(int i=0;i<25;i++) io[i*n+t]=make_uint2(a[i].lo,a[i].hi);
}
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("cuda err %s line %d\n",cudaGetErrorString(e),__LINE__); exit(2);} }while(0)
template<int UU,int NI,int TPB,int MINB> void run(const char* name, uint2* d, int sms){
  int iters=200; int blocks=sms*MINB*4;
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  kk<UU,NI,TPB,MINB><<<blocks,TPB>>>(d,2); CK(cudaDeviceSynchronize());
  float best=1e30f;
  for(int rep=0;rep<3;rep++){ CK(cudaEventRecord(e0)); kk<UU,NI,TPB,MINB><<<blocks,TPB>>>(d,iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms,e0,e1)); if(ms<best)best=ms; }
  double perms=(double)blocks*TPB*iters;
  int nb=0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kk<UU,NI,TPB,MINB>, TPB, 0);
  printf("{\"variant\":\"%s\",\"unroll\":%d,\"imad_rot_lanes\":%d,\"tpb\":%d,\"blocks_per_sm\":%d,\"ms\":%.3f,\"gperm_per_s\":%.3f,\"tera_alg_ops\":%.3f}\n",name,UU,NI,TPB,nb,best,perms/best/1e6,perms*4320/best/1e9);
}
// Occupancy forced by dynamic shared memory: `bps` blocks of 128 threads per SM = `bps` warps per scheduler.  Answers how many
// Keccak warps a scheduler needs to keep its alu pipe full (the fused matrix kernel has 3 resident, ~2 of them in Keccak).
template<int UU> void run_occ(const char* name, uint2* d, int sms, int bps){
  const int TPB=128; int iters=200; int blocks=sms*bps*4;
  size_t smem = (size_t)(220*1024/bps) & ~(size_t)1023; if(smem>200*1024) smem=200*1024;
  CK(cudaFuncSetAttribute(kk<UU,0,TPB,1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  kk<UU,0,TPB,1><<<blocks,TPB,smem>>>(d,2); CK(cudaDeviceSynchronize());
  float best=1e30f;
  for(int rep=0;rep<3;rep++){ CK(cudaEventRecord(e0)); kk<UU,0,TPB,1><<<blocks,TPB,smem>>>(d,iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms,e0,e1)); if(ms<best)best=ms; }
  double perms=(double)blocks*TPB*iters;
  int nb=0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kk<UU,0,TPB,1>, TPB, smem);
  printf("{\"variant\":\"%s\",\"unroll\":%d,\"tpb\":%d,\"blocks_per_sm\":%d,\"warps_per_scheduler\":%d,\"ms\":%.3f,\"gperm_per_s\":%.3f,\"tera_alg_ops\":%.3f}\n",name,UU,TPB,nb,nb,best,perms/best/1e6,perms*4320/best/1e9);
}
int main(int argc, char** argv){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0)); int sms=p.multiProcessorCount;
  uint32_t h[32]; for(int i=0;i<32;i++)h[i]=1u<<i; CK(cudaMemcpyToSymbol(POW2,h,sizeof h));
  uint2* d; size_t n=(size_t)sms*16*1024*25; CK(cudaMalloc(&d,n*sizeof(uint2))); CK(cudaMemset(d,0x5a,n*sizeof(uint2)));
  if(argc>1){  // occupancy study only
    for(int bps=1;bps<=4;bps++){ run_occ<1>("occ_loop1",d,sms,bps); run_occ<2>("occ_loop2",d,sms,bps); }
    return 0;
  }
  run<1,0,256,2>("loop1",d,sms);
  run<2,0,256,2>("loop2",d,sms);
  run<24,0,256,2>("full",d,sms);
  run<1,0,128,4>("loop1_128x4",d,sms);
  run<1,0,128,2>("loop1_128x2",d,sms);
  run<1,0,128,1>("loop1_128x1",d,sms);
  run<1,0,64,1>("loop1_64x1",d,sms);
  run<1,6,256,2>("imad6",d,sms);
  run<1,10,256,2>("imad10",d,sms);
  run<1,14,256,2>("imad14",d,sms);
  run<1,24,256,2>("imad24",d,sms);
  return 0;
}
