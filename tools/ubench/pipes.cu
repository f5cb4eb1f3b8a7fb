// Instruction-throughput microbenchmark for the integer / packed-16x2 / half2 pipes on sm_100a.
// Purpose: size the per-sample instruction budget of the fused SWTPG kernel (DESIGN.md "issue budget").
// Each kernel runs CHAINS independent dependency chains per thread so the pipe, not latency, is measured.
// Output: warp-instructions per clock per SM for every op, and for a few two-op mixes (dual-pipe check).
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <string>
#include <algorithm>
#include <cstring>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

constexpr int CHAINS = 8;
constexpr int ITERS = 8192;
__device__ __forceinline__ unsigned smid() { unsigned r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned long long gtime() { unsigned long long r; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(r)); return r; }
struct Stamp { long long c0, c1; unsigned long long g0, g1; unsigned sm; unsigned pad; };

#define DEFK(name, OPS_PER_ITER, BODY)                                                          \
  __global__ void __launch_bounds__(256) name(unsigned* out, unsigned x, unsigned y, Stamp* cyc) { \
    unsigned r[CHAINS];                                                                          \
    _Pragma("unroll") for (int i = 0; i < CHAINS; ++i) r[i] = threadIdx.x * 2654435761u + i + x; \
    unsigned long long g0 = gtime(); long long t0 = clock64();                                                                    \
    for (int it = 0; it < ITERS; ++it) {                                                         \
      _Pragma("unroll") for (int i = 0; i < CHAINS; ++i) { BODY }                                \
    }                                                                                            \
    long long t1 = clock64(); unsigned long long g1 = gtime();                                                                    \
    unsigned acc = 0;                                                                            \
    _Pragma("unroll") for (int i = 0; i < CHAINS; ++i) acc ^= r[i];                              \
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;                                            \
    if ((threadIdx.x & 31) == 0) { Stamp st; st.c0 = t0; st.c1 = t1; st.g0 = g0; st.g1 = g1; st.sm = smid(); st.pad = 0; cyc[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = st; }                                             \
  }                                                                                              \
  static const int name##_ops = OPS_PER_ITER;

DEFK(k_iadd,   1, asm volatile("add.s32 %0, %0, %1;" : "+r"(r[i]) : "r"(r[(i + 1) % CHAINS]));)
DEFK(k_lop3,   1, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(x), "r"(y));)
DEFK(k_shf,    1, asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(x), "r"(y));)
DEFK(k_prmt,   1, asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(x), "r"(y));)
DEFK(k_imad,   1, asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(x), "r"(y));)
DEFK(k_shl,    1, asm volatile("shl.b32 %0, %0, 1;" : "+r"(r[i]));)
DEFK(k_add16x2,1, asm volatile("add.s16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(y));)
DEFK(k_max16x2,1, asm volatile("max.s16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(r[(i + 1) % CHAINS]));)
DEFK(k_addmax16x2, 1, asm volatile("{.reg .b32 t; add.s16x2 t, %0, %1; max.s16x2 %0, t, %2;}" : "+r"(r[i]) : "r"(x), "r"(y));)
DEFK(k_min3_16x2, 1, asm volatile("{.reg .b32 t; min.s16x2 t, %0, %1; min.s16x2 %0, t, %2;}" : "+r"(r[i]) : "r"(x), "r"(y));)
DEFK(k_hset2,  1, asm volatile("set.gt.u32.f16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(y));)
DEFK(k_hfma2,  1, asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(x), "r"(y));)
DEFK(k_hadd2,  1, asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(y));)
DEFK(k_vimnmx_pred, 2, asm volatile("{.reg .pred p; setp.gt.s32 p, %0, %1; @p add.s32 %0, %0, %2;}" : "+r"(r[i]) : "r"(x), "r"(r[(i + 1) % CHAINS]));)
DEFK(k_dp2a,   1, asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(x), "r"(y));)
DEFK(k_dp4a,   1, asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(x), "r"(y));)
DEFK(k_setp_selp, 2, asm volatile("{.reg .pred p; setp.gt.s32 p, %0, %1; selp.b32 %0, %2, %0, p;}" : "+r"(r[i]) : "r"(x), "r"(y));)
DEFK(k_max32,  1, asm volatile("max.s32 %0, %0, %1;" : "+r"(r[i]) : "r"(r[(i + 1) % CHAINS]));)
DEFK(k_mulhi,  1, asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(x));)
DEFK(k_madhi,  1, asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(x), "r"(y));)
DEFK(k_mix_mulhi_lop, 2, asm volatile("mul.hi.u32 %0, %0, %1; lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(x), "r"(y));)
DEFK(k_mix_mulhi_imad, 2, asm volatile("mul.hi.u32 %0, %0, %1; mad.lo.s32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(x), "r"(y));)
DEFK(k_mulwide, 1, asm volatile("{.reg .b64 t; mul.wide.u32 t, %0, %1; cvt.u32.u64 %0, t;}" : "+r"(r[i]) : "r"(x));)
DEFK(k_mulwide_hi, 1, asm volatile("{.reg .b64 t; .reg .b32 lo; mul.wide.u32 t, %0, %1; mov.b64 {lo, %0}, t;}" : "+r"(r[i]) : "r"(x));)
DEFK(k_shr,    1, asm volatile("shr.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(y));)
DEFK(k_bfe,    1, asm volatile("bfe.u32 %0, %0, %1, 14;" : "+r"(r[i]) : "r"(y));)
DEFK(k_mix_lds_lop, 2, { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"((r[i] & 0xFFCu))); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(v), "r"(y)); })
DEFK(k_ffma,   1, asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float*)&r[i]) : "f"(1.0001f), "f"(0.5f));)
// two-op mixes: is issue shared or are there two independent pipes?
DEFK(k_mix_lop_imad, 2, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96; mad.lo.s32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(x), "r"(y));)
DEFK(k_mix_add16_hfma2, 2, asm volatile("add.s16x2 %0, %0, %2; fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(x), "r"(y));)
DEFK(k_mix_add16_lop, 2, asm volatile("add.s16x2 %0, %0, %2; lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(x), "r"(y));)
DEFK(k_mix_max16_imad, 2, asm volatile("max.s16x2 %0, %0, %2; mad.lo.s32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(x), "r"(y));)
DEFK(k_mix_max16_lop, 2, asm volatile("max.s16x2 %0, %0, %2; lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(x), "r"(y));)
DEFK(k_mix_hset2_lop, 2, asm volatile("set.gt.u32.f16x2 %0, %0, %2; lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(x), "r"(y));)
DEFK(k_mix_hset2_hfma2, 2, asm volatile("set.gt.u32.f16x2 %0, %0, %2; fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(x), "r"(y));)
DEFK(k_mix_shf_imad, 2, asm volatile("shf.r.wrap.b32 %0, %0, %1, %2; mad.lo.s32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(x), "r"(y));)
DEFK(k_mix_iadd_imad, 2, asm volatile("add.s32 %0, %0, %2; mad.lo.s32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(x), "r"(y));)
DEFK(k_mix_add16_max16, 2, asm volatile("add.s16x2 %0, %0, %2; max.s16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(x), "r"(y));)

// shared-memory load throughput (conflict-free 32-bit and 128-bit, and broadcast-heavy pattern like the 14-bit row read)
__global__ void __launch_bounds__(256) k_lds32(unsigned* out, unsigned x, unsigned y, Stamp* cyc) {
  __shared__ unsigned sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * x;
  __syncthreads();
  unsigned acc = 0; unsigned idx = threadIdx.x;
  unsigned long long g0 = gtime(); long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { acc += sm[(idx + i * 256 + it * 32) & 4095]; }
  }
  long long t1 = clock64(); unsigned long long g1 = gtime();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + y;
  if ((threadIdx.x & 31) == 0) { Stamp st; st.c0 = t0; st.c1 = t1; st.g0 = g0; st.g1 = g1; st.sm = smid(); st.pad = 0; cyc[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = st; }
}
static const int k_lds32_ops = 1;
__global__ void __launch_bounds__(256) k_lds128(unsigned* out, unsigned x, unsigned y, Stamp* cyc) {
  __shared__ uint4 sm[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_uint4(i * x, i, x, y);
  __syncthreads();
  unsigned acc = 0; unsigned idx = threadIdx.x;
  unsigned long long g0 = gtime(); long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { uint4 v = sm[(idx + i * 256 + it * 32) & 2047]; acc += v.x ^ v.y ^ v.z ^ v.w; }
  }
  long long t1 = clock64(); unsigned long long g1 = gtime();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + y;
  if ((threadIdx.x & 31) == 0) { Stamp st; st.c0 = t0; st.c1 = t1; st.g0 = g0; st.g1 = g1; st.sm = smid(); st.pad = 0; cyc[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = st; }
}
static const int k_lds128_ops = 1;
// 14-bit row pattern: lane l reads word (14*2*l)>>5 and the next one (pairs of channels) from a 112-byte row
__global__ void __launch_bounds__(256) k_lds_row14(unsigned* out, unsigned x, unsigned y, Stamp* cyc) {
  __shared__ unsigned sm[4096 + 32];
  for (int i = threadIdx.x; i < 4096 + 32; i += blockDim.x) sm[i] = i * x;
  __syncthreads();
  unsigned acc = 0; unsigned lane = threadIdx.x & 31; unsigned w = (28 * lane) >> 5; unsigned warp = threadIdx.x >> 5;
  unsigned long long g0 = gtime(); long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS / 2; ++i) {
      unsigned base = ((it * 4 + i) * 28 + warp * 7) & 4095;
      acc += __funnelshift_r(sm[base + w], sm[base + w + 1], lane * 28);
    }
  }
  long long t1 = clock64(); unsigned long long g1 = gtime();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + y;
  if ((threadIdx.x & 31) == 0) { Stamp st; st.c0 = t0; st.c1 = t1; st.g0 = g0; st.g1 = g1; st.sm = smid(); st.pad = 0; cyc[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = st; }
}
static const int k_lds_row14_ops = 1;  // counts LDS only: CHAINS loads per iteration (2 per pair)

typedef void (*kern_t)(unsigned*, unsigned, unsigned, Stamp*);
struct Entry { const char* name; kern_t k; int ops; };

int main() {
  cudaDeviceProp prop; CHECK(cudaGetDeviceProperties(&prop, 0));
  int nsm = prop.multiProcessorCount;
  printf("device=%s sms=%d clock_khz=%d asyncEngines=%d\n", prop.name, nsm, prop.clockRate, prop.asyncEngineCount);
  const int ctas_per_sm = 4, threads = 256;
  int grid = nsm * ctas_per_sm;
  unsigned* out; Stamp* cyc;
  CHECK(cudaMalloc(&out, sizeof(unsigned) * grid * threads));
  int nw = grid * threads / 32;
  CHECK(cudaMalloc(&cyc, sizeof(Stamp) * nw));
  std::vector<Stamp> h(nw);
  { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a); for (int i = 0; i < 400; ++i) k_ffma<<<grid, threads>>>(out, 3, 5, cyc); cudaEventRecord(b); CHECK(cudaDeviceSynchronize()); float ms; cudaEventElapsedTime(&ms, a, b); printf("clock warm-up: %.1f ms\n", ms); }
#define E(n) { #n, n, n##_ops }
  Entry es[] = { E(k_iadd), E(k_lop3), E(k_shf), E(k_prmt), E(k_imad), E(k_shl), E(k_add16x2), E(k_max16x2), E(k_addmax16x2),
                 E(k_min3_16x2), E(k_hset2), E(k_hfma2), E(k_hadd2), E(k_vimnmx_pred), E(k_dp2a), E(k_dp4a), E(k_setp_selp), E(k_max32), E(k_mulhi), E(k_madhi), E(k_mix_mulhi_lop), E(k_mix_mulhi_imad), E(k_mulwide), E(k_mulwide_hi), E(k_shr), E(k_bfe), E(k_ffma),
                 E(k_mix_lop_imad), E(k_mix_add16_hfma2), E(k_mix_add16_lop), E(k_mix_max16_imad), E(k_mix_max16_lop),
                 E(k_mix_hset2_lop), E(k_mix_hset2_hfma2), E(k_mix_shf_imad), E(k_mix_iadd_imad), E(k_mix_add16_max16),
                 E(k_lds32), E(k_lds128), E(k_lds_row14) };
  printf("%-22s %10s %12s %14s\n", "kernel", "ms", "cycles(avg)", "warpinstr/clk/SM");
  for (auto& e : es) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    e.k<<<grid, threads>>>(out, 3, 5, cyc);  // warm-up
    CHECK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    e.k<<<grid, threads>>>(out, 3, 5, cyc);
    cudaEventRecord(b);
    CHECK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, a, b);
    CHECK(cudaMemcpy(h.data(), cyc, sizeof(Stamp) * nw, cudaMemcpyDeviceToHost));
    // per-SM span: max(c1) - min(c0) over the warps that ran on that SM (clock64 is per-SM)
    std::vector<long long> c0(256, (long long)1 << 62), c1(256, 0);
    unsigned long long g0 = ~0ull, g1 = 0;
    for (auto& st : h) { c0[st.sm] = std::min(c0[st.sm], st.c0); c1[st.sm] = std::max(c1[st.sm], st.c1); g0 = std::min(g0, st.g0); g1 = std::max(g1, st.g1); }
    double avg = 0; int n = 0; for (int i = 0; i < 256; ++i) if (c1[i]) { avg += double(c1[i] - c0[i]); ++n; } avg /= n;
    double winstr = double(ctas_per_sm) * (threads / 32) * ITERS * CHAINS * e.ops;
    printf("%-22s %10.3f %12.0f %14.3f   sm_clk=%.0f MHz\n", e.name, ms, avg, winstr / avg, avg / double(g1 - g0) * 1e3);
  }
  // H2D / D2H pinned bandwidth (ingest roofline denominator)
  size_t bytes = size_t(1) << 30;
  void *hbuf, *dbuf; CHECK(cudaMallocHost(&hbuf, bytes)); CHECK(cudaMalloc(&dbuf, bytes));
  memset(hbuf, 1, bytes);
  cudaStream_t s[4]; for (auto& st : s) cudaStreamCreate(&st);
  for (int nstream : {1, 2, 4}) {
    for (int dir = 0; dir < 2; ++dir) {
      float best = 1e30f;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        CHECK(cudaDeviceSynchronize());
        cudaEventRecord(a, 0);
        size_t chunk = bytes / nstream;
        for (int i = 0; i < nstream; ++i) {
          cudaStreamWaitEvent(s[i], a, 0);
          if (dir == 0) cudaMemcpyAsync((char*)dbuf + i * chunk, (char*)hbuf + i * chunk, chunk, cudaMemcpyHostToDevice, s[i]);
          else cudaMemcpyAsync((char*)hbuf + i * chunk, (char*)dbuf + i * chunk, chunk, cudaMemcpyDeviceToHost, s[i]);
        }
        for (int i = 0; i < nstream; ++i) { cudaEvent_t d; cudaEventCreate(&d); cudaEventRecord(d, s[i]); cudaStreamWaitEvent(0, d, 0); }
        cudaEventRecord(b, 0);
        CHECK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
      }
      printf("%s pinned 1GiB streams=%d: %.2f GB/s\n", dir == 0 ? "H2D" : "D2H", nstream, bytes / best / 1e6);
    }
  }
  return 0;
}
