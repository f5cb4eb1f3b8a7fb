// Dependent-issue latency of the instructions the SWTPG tick loop is made of (sm_100a): ONE warp runs a chain of N dependent
// ops; cycles per op = (clock64 delta) / N. Also the frugal-pedestal recurrence exactly as the kernel issues it, alone in a
// warp, to see what a lone link can reach (DESIGN.md "Under-filled GPUs").
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/bin/latency tools/ubench/latency.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

constexpr int N = 4096;

#define LATK(name, BODY)                                                                   \
  __global__ void name(uint32_t* out, long long* cyc, uint32_t x, uint32_t y)              \
  {                                                                                        \
    uint32_t r = threadIdx.x + x, q = y;                                                   \
    long long t0 = clock64();                                                              \
    _Pragma("unroll 16") for (int i = 0; i < N; ++i) { BODY }                              \
    long long t1 = clock64();                                                              \
    out[threadIdx.x] = r + q;                                                              \
    if (threadIdx.x == 0)                                                                  \
      *cyc = t1 - t0;                                                                      \
  }

LATK(l_iadd, asm volatile("add.u32 %0, %0, %1;" : "+r"(r) : "r"(y));)
LATK(l_lop3, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r) : "r"(x), "r"(y));)
LATK(l_shf, asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(r) : "r"(x), "r"(y));)
LATK(l_imad, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r) : "r"(x), "r"(y));)
LATK(l_add16x2, asm volatile("add.s16x2 %0, %0, %1;" : "+r"(r) : "r"(y));)
LATK(l_max16x2, asm volatile("max.s16x2 %0, %0, %1;" : "+r"(r) : "r"(y));)
LATK(l_addmax16x2, asm volatile("{.reg .b32 t; add.s16x2 t, %0, %1; max.s16x2 %0, t, %2;}" : "+r"(r) : "r"(x), "r"(y));)
LATK(l_viaddmin_relu, r = __viaddmin_s16x2_relu(r, x, y);)
LATK(l_vimax3, r = __vimax3_s16x2(r, x, y);)
LATK(l_hadd2, asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(r) : "r"(y));)
LATK(l_hfma2, asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(r) : "r"(x), "r"(y));)
LATK(l_hset2, asm volatile("set.eq.u32.f16x2 %0, %0, %1;" : "+r"(r) : "r"(y));)
LATK(l_hset2_bf, asm volatile("set.ne.f16x2.f16x2 %0, %0, %1;" : "+r"(r) : "r"(y));)
// cross-pipe hops
LATK(l_add16_hfma2, asm volatile("add.s16x2 %0, %0, %1;" : "+r"(r) : "r"(y)); asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(r) : "r"(x), "r"(y));)
LATK(l_add16_hset2, asm volatile("add.s16x2 %0, %0, %1;" : "+r"(r) : "r"(y)); asm volatile("set.eq.u32.f16x2 %0, %0, %1;" : "+r"(r) : "r"(y));)
LATK(l_hfma2_hset2, asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(r) : "r"(x), "r"(y)); asm volatile("set.eq.u32.f16x2 %0, %0, %1;" : "+r"(r) : "r"(y));)
LATK(l_add16_lop3, asm volatile("add.s16x2 %0, %0, %1;" : "+r"(r) : "r"(y)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r) : "r"(x), "r"(y));)
LATK(l_iadd_imad, asm volatile("add.u32 %0, %0, %1;" : "+r"(r) : "r"(y)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r) : "r"(x), "r"(y));)

// the pedestal recurrence of the kernel (PackedSimpleWibEth::pedestal_step), fp16-subnormal accumulator form: per tick
//   sg1 = clamp(S + Mq, 0, 2); T = A + sg1; upm = (T == cUp); dn1 = sat(-T - L); keep = |T| != cUp; A = keep*T - tiny; Mq += upm + dn1
__global__ void l_pedestal(uint32_t* out, long long* cyc, uint32_t x, uint32_t y)
{
  uint32_t Mq = threadIdx.x + x, A = y, S = x ^ 0x12345678u;
  const uint32_t cUp = 0x000B000Bu, cDn = 0x800A800Au;
  long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) {
    const uint32_t sg1 = __viaddmin_s16x2_relu(S, Mq, 0x00020002u);
    uint32_t T, upm, dn1, keep;
    asm volatile("add.rn.f16x2 %0, %1, %2;" : "=r"(T) : "r"(A), "r"(sg1));
    asm volatile("set.eq.u32.f16x2 %0, %1, %2;" : "=r"(upm) : "r"(T), "r"(cUp));
    asm volatile("fma.rn.sat.f16x2 %0, %1, %2, %3;" : "=r"(dn1) : "r"(T), "r"(0xBC00BC00u), "r"(cDn));
    asm volatile("{.reg .b32 t; abs.f16x2 t, %1; set.ne.f16x2.f16x2 %0, t, %2;}" : "=r"(keep) : "r"(T), "r"(cUp));
    asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(A) : "r"(keep), "r"(T), "r"(0x80018001u));
    asm volatile("add.s16x2 %0, %0, %1;" : "+r"(Mq) : "r"(upm));
    asm volatile("add.s16x2 %0, %0, %1;" : "+r"(Mq) : "r"(dn1));
    S += 0x00030005u; // stand-in for the next sample (independent of the chain)
  }
  long long t1 = clock64();
  out[threadIdx.x] = Mq + A;
  if (threadIdx.x == 0)
    *cyc = t1 - t0;
}
// integer form of the same recurrence (SWTPG_FLOAT_ACC=0): T = A + sg1; up = max(T - (L+1) .. ); A = (T-1) & ~(upm|dn); Mq += ...
__global__ void l_pedestal_int(uint32_t* out, long long* cyc, uint32_t x, uint32_t y)
{
  uint32_t Mq = threadIdx.x + x, A = y, S = x ^ 0x12345678u;
  const uint32_t cUp = 0xFFF5FFF5u, cDn = 0x00090009u;
  long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) {
    const uint32_t sg1 = __viaddmin_s16x2_relu(S, Mq, 0x00020002u);
    uint32_t T, up, dn;
    asm volatile("add.s16x2 %0, %1, %2;" : "=r"(T) : "r"(A), "r"(sg1));
    asm volatile("{.reg .b32 t; add.s16x2 t, %1, %2; max.s16x2 %0, t, %3;}" : "=r"(up) : "r"(T), "r"(cUp), "r"(0u));
    asm volatile("{.reg .b32 t; add.s16x2 t, %1, %2; min.s16x2 %0, t, %3;}" : "=r"(dn) : "r"(T), "r"(cDn), "r"(0u));
    const uint32_t upm = up * 0xFFFFu;
    uint32_t Tm;
    asm volatile("add.s16x2 %0, %1, %2;" : "=r"(Tm) : "r"(T), "r"(0xFFFFFFFFu));
    A = Tm & ~(upm | dn);
    asm volatile("add.s16x2 %0, %0, %1;" : "+r"(Mq) : "r"(upm | (dn & 0x00010001u)));
    S += 0x00030005u;
  }
  long long t1 = clock64();
  out[threadIdx.x] = Mq + A;
  if (threadIdx.x == 0)
    *cyc = t1 - t0;
}

int main()
{
  CK(cudaSetDevice(0));
  uint32_t* d_out;
  long long* d_cyc;
  CK(cudaMalloc(&d_out, 4096));
  CK(cudaMalloc(&d_cyc, 8));
#define RUN(k, ops)                                                                        \
  {                                                                                        \
    k<<<1, 32>>>(d_out, d_cyc, 3u, 5u);                                                    \
    k<<<1, 32>>>(d_out, d_cyc, 3u, 5u);                                                    \
    CK(cudaDeviceSynchronize());                                                           \
    long long c = 0;                                                                       \
    CK(cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost));                                  \
    printf("%-18s %7.2f cycles per %s\n", #k, double(c) / N / (ops), (ops) == 1 ? "op" : "op (avg over the pair)"); \
  }
  RUN(l_iadd, 1) RUN(l_lop3, 1) RUN(l_shf, 1) RUN(l_imad, 1) RUN(l_add16x2, 1) RUN(l_max16x2, 1) RUN(l_addmax16x2, 1) RUN(l_viaddmin_relu, 1)
  RUN(l_vimax3, 1) RUN(l_hadd2, 1) RUN(l_hfma2, 1) RUN(l_hset2, 1) RUN(l_hset2_bf, 1)
  RUN(l_add16_hfma2, 2) RUN(l_add16_hset2, 2) RUN(l_hfma2_hset2, 2) RUN(l_add16_lop3, 2) RUN(l_iadd_imad, 2)
  {
    l_pedestal<<<1, 32>>>(d_out, d_cyc, 3u, 5u);
    l_pedestal<<<1, 32>>>(d_out, d_cyc, 3u, 5u);
    CK(cudaDeviceSynchronize());
    long long c = 0;
    CK(cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost));
    printf("pedestal recurrence (fp16 accumulator form): %.2f cycles per tick, one warp alone\n", double(c) / N);
    l_pedestal_int<<<1, 32>>>(d_out, d_cyc, 3u, 5u);
    l_pedestal_int<<<1, 32>>>(d_out, d_cyc, 3u, 5u);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost));
    printf("pedestal recurrence (integer accumulator form): %.2f cycles per tick, one warp alone\n", double(c) / N);
  }
  return 0;
}
