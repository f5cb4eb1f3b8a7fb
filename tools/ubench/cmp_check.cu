// Exhaustive check of the two "compare through the float comparator" tricks the kernels rely on (run on the GPU):
//   gt2_mask_nonneg : set.gt.u32.f16x2   for a in [-16384, 16384], b in [0, 0x7BFF]
//   gt2_mask_bf16   : set.gtu.u32.bf16x2 for a in [0, 32767],     b in [0, 32640]
// and of exact integer arithmetic on fp16 subnormals (add / fma.sat / set.eq) for |v| <= 1023.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ unsigned long long g_bad[4];
__global__ void check_bf16()
{
  const uint32_t a = blockIdx.x;                      // 0..32767
  for (uint32_t b = threadIdx.x; b <= 32640; b += blockDim.x) {
    uint32_t r, x = a | (a << 16), y = b | (b << 16);
    asm("set.gtu.u32.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(y));
    const uint32_t want = a > b ? 0xFFFFFFFFu : 0u;
    if (r != want) atomicAdd(&g_bad[0], 1ull);
  }
}
__global__ void check_f16()
{
  const int a = int(blockIdx.x) - 16384;              // -16384..16384
  for (uint32_t b = threadIdx.x; b <= 0x7BFF; b += blockDim.x) {
    uint32_t r, x = (uint32_t(a) & 0xFFFF) | (uint32_t(a) << 16), y = b | (b << 16);
    asm("set.gt.u32.f16x2 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(y));
    const uint32_t want = a > int(b) ? 0xFFFFFFFFu : 0u;
    if (r != want) atomicAdd(&g_bad[1], 1ull);
  }
}
__device__ uint32_t sm(int v) { return v < 0 ? (0x8000u | uint32_t(-v)) : uint32_t(v); }
__global__ void check_subnormal()
{
  const int a = int(blockIdx.x) - 1023;               // -1023..1023
  for (int b = int(threadIdx.x) - 1023; b <= 1023; b += int(blockDim.x)) {
    if (a + b > 1023 || a + b < -1023) continue;
    uint32_t r, x = sm(a) * 0x10001u, y = sm(b) * 0x10001u;
    asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(y));
    uint32_t want = sm(a + b) * 0x10001u;
    if (r != want && !(a + b == 0 && (r & 0x7FFF7FFFu) == 0)) atomicAdd(&g_bad[2], 1ull);
    asm("set.eq.u32.f16x2 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(y));
    if (r != (a == b ? 0xFFFFFFFFu : 0u)) atomicAdd(&g_bad[3], 1ull);
  }
}
int main()
{
  check_bf16<<<32768, 256>>>();
  check_f16<<<32769, 256>>>();
  check_subnormal<<<2047, 256>>>();
  cudaDeviceSynchronize();
  unsigned long long bad[4];
  cudaMemcpyFromSymbol(bad, g_bad, sizeof bad);
  printf("bf16 gtu mismatches %llu, f16 gt mismatches %llu, subnormal add mismatches %llu, subnormal eq mismatches %llu\n", bad[0], bad[1], bad[2], bad[3]);
  return (bad[0] | bad[1] | bad[2] | bad[3]) ? 1 : 0;
}
