// Design probe for the streaming ingest (DESIGN.md "Ingest"): how fast can SM-issued copies pull 7200-byte frames out of
// pinned, mapped HOST memory into HBM, against the copy engine's plain cudaMemcpyAsync? Two forms:
//   tma : one thread per CTA runs a ring of cp.async.bulk global(host)->shared, then shared->global(HBM)
//   lsu : one warp per frame, 16-byte loads (several in flight per lane) and stores
// Source frames are addressed through a per-frame pointer table (scattered latency-buffer slots), as the product does.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_probe tools/ubench/gather_probe.cu
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <random>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct Item { const uint8_t* src; uint64_t dst_off; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

template<int STAGES>
__global__ void __launch_bounds__(32) gather_tma(const Item* items, uint32_t n, uint8_t* dst, uint32_t bytes)
{
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + size_t(STAGES) * bytes);
  if (threadIdx.x != 0)
    return;
  for (int s = 0; s < STAGES; ++s)
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[s])));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  const uint32_t first = blockIdx.x, step = gridDim.x;
  const uint32_t mine = first < n ? (n - first + step - 1) / step : 0;
  auto load = [&](uint32_t k) { // k-th item of this CTA
    const uint32_t s = k % STAGES;
    const Item it = items[first + k * step];
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + size_t(s) * bytes)),
                 "l"(it.src), "r"(bytes), "r"(smem_u32(&bar[s]))
                 : "memory");
  };
  for (uint32_t k = 0; k < mine && k < STAGES - 1; ++k)
    load(k);
  for (uint32_t k = 0; k < mine; ++k) {
    const uint32_t s = k % STAGES, parity = (k / STAGES) & 1u;
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}" ::"r"(smem_u32(&bar[s])), "r"(parity) : "memory");
    const Item it = items[first + k * step];
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + it.dst_off), "r"(smem_u32(smem + size_t(s) * bytes)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); // the store of item k-1 has read its stage: refill it
    if (k + STAGES - 1 < mine)
      load(k + STAGES - 1);
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template<int UNROLL>
__global__ void __launch_bounds__(256) gather_lsu(const Item* items, uint32_t n, uint8_t* dst, uint32_t vecs)
{
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u, warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t i = warp; i < n; i += warps) {
    const Item it = items[i];
    const uint4* s = reinterpret_cast<const uint4*>(it.src);
    uint4* d = reinterpret_cast<uint4*>(dst + it.dst_off);
    for (uint32_t j0 = lane; j0 < vecs; j0 += 32u * UNROLL) {
      uint4 v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
        if (j0 + 32u * u < vecs)
          asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(s + j0 + 32u * u));
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
        if (j0 + 32u * u < vecs)
          d[j0 + 32u * u] = v[u];
    }
  }
}

int main(int argc, char** argv)
{
  const uint32_t bytes = argc > 1 ? atoi(argv[1]) : 7200;
  const uint32_t n = argc > 2 ? atoi(argv[2]) : 240 * 64 * 8; // units per launch
  const bool shuffle = argc > 3 ? atoi(argv[3]) != 0 : true;
  CK(cudaSetDevice(0));
  uint8_t *h = nullptr, *d = nullptr;
  const size_t total = size_t(bytes) * n;
  CK(cudaHostAlloc(&h, total, cudaHostAllocMapped | cudaHostAllocPortable));
  CK(cudaMalloc(&d, total));
  for (size_t i = 0; i < total; i += 8)
    *reinterpret_cast<uint64_t*>(h + i) = i * 0x9E3779B97F4A7C15ull;
  uint8_t* hd = nullptr;
  CK(cudaHostGetDevicePointer(&hd, h, 0));
  std::vector<uint32_t> order(n);
  std::iota(order.begin(), order.end(), 0u);
  if (shuffle) { // frames of one link are contiguous in its latency buffer; links interleave
    std::mt19937 rng(1);
    const uint32_t run = 64;
    std::vector<uint32_t> runs(n / run);
    std::iota(runs.begin(), runs.end(), 0u);
    std::shuffle(runs.begin(), runs.end(), rng);
    for (uint32_t r = 0; r < n / run; ++r)
      for (uint32_t j = 0; j < run; ++j)
        order[r * run + j] = runs[r] * run + j;
  }
  std::vector<Item> items(n);
  for (uint32_t i = 0; i < n; ++i)
    items[i] = { hd + size_t(order[i]) * bytes, uint64_t(i) * bytes };
  Item* d_items = nullptr;
  CK(cudaMalloc(&d_items, n * sizeof(Item)));
  CK(cudaMemcpy(d_items, items.data(), n * sizeof(Item), cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  std::vector<uint8_t> back(total);
  auto check = [&](const char* what) {
    CK(cudaMemcpy(back.data(), d, total, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (uint32_t i = 0; i < n && bad == 0; ++i)
      if (memcmp(back.data() + size_t(i) * bytes, h + size_t(order[i]) * bytes, bytes) != 0)
        ++bad;
    printf("  %s check: %s\n", what, bad ? "MISMATCH" : "ok");
    CK(cudaMemset(d, 0, total));
  };
  auto timeit = [&](const char* name, auto&& fn) {
    fn();
    CK(cudaDeviceSynchronize());
    float best = 1e9f;
    for (int r = 0; r < 3; ++r) {
      CK(cudaEventRecord(e0));
      fn();
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      best = std::min(best, ms);
    }
    CK(cudaGetLastError());
    printf("%-28s %8.3f ms  %6.2f GB/s\n", name, best, total / (best * 1e-3) / 1e9);
    fflush(stdout);
  };
  printf("unit %u B x %u units = %.1f MB, %s order\n", bytes, n, total / 1e6, shuffle ? "run-shuffled" : "linear");
  timeit("cudaMemcpyAsync (1 copy)", [&] { CK(cudaMemcpyAsync(d, h, total, cudaMemcpyHostToDevice, 0)); });
  timeit("cudaMemcpyAsync (per 64 units)", [&] {
    for (uint32_t i = 0; i < n; i += 64)
      CK(cudaMemcpyAsync(d + size_t(i) * bytes, h + size_t(order[i]) * bytes, size_t(64) * bytes, cudaMemcpyHostToDevice, 0));
  });
  check("memcpy");
  { // the same per-run copies dealt round-robin over K streams: do the copy engines overlap the per-copy overhead?
    cudaStream_t st[16];
    for (int i = 0; i < 16; ++i)
      CK(cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking));
    for (int K : { 2, 4, 8, 16 }) {
      char nm[64];
      snprintf(nm, sizeof nm, "memcpy per 64 units, %d streams", K);
      timeit(nm, [&] {
        cudaEvent_t j[16];
        for (int i = 0; i < K; ++i) { // fork from the timing stream
          CK(cudaEventCreateWithFlags(&j[i], cudaEventDisableTiming));
          CK(cudaEventRecord(j[i], 0));
          CK(cudaStreamWaitEvent(st[i], j[i], 0));
        }
        uint32_t k = 0;
        for (uint32_t i = 0; i < n; i += 64, ++k)
          CK(cudaMemcpyAsync(d + size_t(i) * bytes, h + size_t(order[i]) * bytes, size_t(64) * bytes, cudaMemcpyHostToDevice, st[k % K]));
        for (int i = 0; i < K; ++i) { // join
          CK(cudaEventRecord(j[i], st[i]));
          CK(cudaStreamWaitEvent(0, j[i], 0));
          CK(cudaEventDestroy(j[i]));
        }
      });
    }
    check("memcpy streams");
  }
#define TMA(ST, G)                                                                                                                  \
  {                                                                                                                                 \
    const size_t sm = size_t(ST) * bytes + ST * 8;                                                                                  \
    if (sm <= 227 * 1024) {                                                                                                         \
    CK(cudaFuncSetAttribute(gather_tma<ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm)));                                 \
    char nm[64];                                                                                                                    \
    snprintf(nm, sizeof nm, "tma stages=%d ctas=%d", ST, G);                                                                        \
    timeit(nm, [&] { gather_tma<ST><<<G, 32, sm>>>(d_items, n, d, bytes); });                                                       \
    }                                                                                                                               \
  }
  TMA(2, 16) TMA(4, 16) TMA(8, 16) TMA(4, 32) TMA(8, 32) TMA(4, 64) TMA(8, 64) TMA(16, 64) TMA(4, 148) TMA(8, 148) TMA(8, 296) TMA(24, 148)
  check("tma");
#define LSU(UN, G)                                                                                                                  \
  {                                                                                                                                 \
    char nm[64];                                                                                                                    \
    snprintf(nm, sizeof nm, "lsu unroll=%d ctas=%d", UN, G);                                                                        \
    timeit(nm, [&] { gather_lsu<UN><<<G, 256>>>(d_items, n, d, bytes / 16); });                                                     \
  }
  LSU(4, 16) LSU(8, 16) LSU(4, 64) LSU(8, 64) LSU(15, 64) LSU(4, 148) LSU(8, 148) LSU(15, 148) LSU(8, 592)
  check("lsu");
  return 0;
}
