// Device-side ordering of a batch's TP list (SWTPG_FLAG_SORTED_TPS): the records the fused kernels append through a global
// cursor come out in whatever order the warps flushed them; downstream (TriggerPrimitiveTypeAdapter::operator<,
// include/fdreadoutlibs/TriggerPrimitiveTypeAdapter.hpp:26-29; the skip list of src/TPCTPRequestHandler.cpp:99-193) wants
// (time_start, link, channel). Ordering a 64-frame batch of 148 APAs on one host core takes 17-48 ms — 30-70x the fused kernel's
// time — so the list is ordered where it lies, in HBM, before it crosses the host link:
//
//   1. tp_minmax_kernel    range of time_start and the bits in which the values differ -> host (32 bytes): key width
//   2. tp_keys_kernel      key = (time_start - min) << (link bits + channel bits) | link << channel bits | channel, value = index
//   3. per 8-bit digit of the key, least significant first (LSD radix sort, stable):
//        radix_hist_kernel     one WARP per tile of consecutive records: digit histogram in shared memory -> hist[digit][tile]
//        radix_scan_kernel     one CTA per digit: keys with a smaller digit (digit totals) + keys of that digit in earlier tiles
//        radix_scatter_kernel  one warp per tile again: 32 records per round, rank among equal digits by __match_any_sync,
//                              running per-digit bases in shared memory; records of a tile keep their order (stability)
//   4. tp_gather_kernel    sorted[i] = tps[value[i]] (two 16-byte loads / stores per record) + a flag if two neighbours share a key
//
// 37-40 key bits for a bench batch = 5 digits = 18 small launches; everything is sized from n, which the host knows at that
// point (it has just read the batch's TP count). Equal keys (cannot come out of one handle unless a link's timestamps jump
// backwards) keep emission order here; the caller then finishes with the host's tie-break (swtpg_sort_tps) so that both
// orderings are identical in every case. Keys wider than 64 bits (links with unrelated timestamps) are left to the host too.
#include "swtpg_handle.h"

#include <algorithm>

namespace swtpg_internal {

namespace {

constexpr int kDigitBits = 8, kBuckets = 1 << kDigitBits;
constexpr int kWarpsPerCta = 4;
constexpr uint32_t kMaxTiles = 4096;

__global__ void
tp_minmax_kernel(const swtpg_tp* __restrict__ tps, uint32_t n, unsigned long long* __restrict__ mm)
{
  // mm = {min, max, tie flag (set later), OR of (time_start XOR time_start of record 0)}: bits that are the same in every
  // time_start — the low five always are when frame timestamps are multiples of 32 — are the same in (t - min) too and need no digit
  unsigned long long lo = ~0ull, hi = 0ull, diff = 0ull;
  const unsigned long long t0 = tps[0].time_start;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned long long t = tps[i].time_start;
    lo = min(lo, t);
    hi = max(hi, t);
    diff |= t ^ t0;
  }
  for (int o = 16; o; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xFFFFFFFFu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xFFFFFFFFu, hi, o));
    diff |= __shfl_xor_sync(0xFFFFFFFFu, diff, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(mm, lo);
    atomicMax(mm + 1, hi);
    if (diff)
      atomicOr(mm + 3, diff);
  }
}

__global__ void
tp_keys_kernel(const swtpg_tp* __restrict__ tps, uint32_t n, unsigned long long tmin, uint32_t time_shift, uint32_t link_bits,
               uint32_t channel_bits, unsigned long long* __restrict__ keys, uint32_t* __restrict__ idx)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  // one 16-byte load for {adc_peak, channel, link} would drag the whole record in anyway: read the two fields' sectors
  const swtpg_tp& r = tps[i];
  const unsigned long long t = (r.time_start - tmin) >> time_shift;
  const unsigned sh = link_bits + channel_bits;
  keys[i] = (sh < 64 ? t << sh : 0ull) | (static_cast<unsigned long long>(r.link) << channel_bits) | r.channel;
  idx[i] = i;
}

__device__ __forceinline__ uint32_t
digit_of(unsigned long long key, uint32_t shift)
{
  return static_cast<uint32_t>(key >> shift) & (kBuckets - 1);
}

// hist[d * n_tiles + tile] = number of keys of `tile` whose current digit is d
__global__ void __launch_bounds__(32 * kWarpsPerCta)
radix_hist_kernel(const unsigned long long* __restrict__ keys, uint32_t n, uint32_t shift, uint32_t tile_elems, uint32_t n_tiles,
                  uint32_t* __restrict__ hist, uint32_t* __restrict__ totals)
{
  __shared__ uint32_t cnt[kWarpsPerCta][kBuckets];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tile = blockIdx.x * kWarpsPerCta + warp;
  for (int d = lane; d < kBuckets; d += 32)
    cnt[warp][d] = 0;
  __syncwarp();
  if (tile >= n_tiles)
    return;
  const uint32_t begin = tile * tile_elems, end = min(n, begin + tile_elems);
  for (uint32_t i = begin + lane; i < end; i += 32)
    atomicAdd(&cnt[warp][digit_of(keys[i], shift)], 1u);
  __syncwarp();
  for (int d = lane; d < kBuckets; d += 32) {
    const uint32_t c = cnt[warp][d];
    hist[size_t(d) * n_tiles + tile] = c;
    if (c)
      atomicAdd(&totals[d], c);
  }
}

// One CTA per digit d: hist[d][0..n_tiles) becomes the position of each tile's first key with that digit = (keys with a smaller
// digit, from the digit totals the histogram pass accumulated) + (keys with digit d in earlier tiles). n_tiles <= kMaxTiles, so a
// thread owns at most kMaxTiles / 256 consecutive tiles and one block scan does it.
__global__ void __launch_bounds__(256)
radix_scan_kernel(uint32_t* __restrict__ hist, const uint32_t* __restrict__ totals, uint32_t n_tiles)
{
  __shared__ uint32_t warp_sums[8];
  __shared__ uint32_t below;
  const uint32_t d = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t v = tid < d ? totals[tid] : 0u;
  for (int o = 16; o; o >>= 1)
    v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  if (lane == 0)
    warp_sums[warp] = v;
  __syncthreads();
  if (tid == 0) {
    uint32_t t = 0;
    for (int w = 0; w < 8; ++w)
      t += warp_sums[w];
    below = t;
  }
  __syncthreads();
  const uint32_t base = below;
  uint32_t* row = hist + size_t(d) * n_tiles;
  constexpr uint32_t kPer = kMaxTiles / 256;
  const uint32_t per = (n_tiles + 255) / 256;
  const uint32_t begin = min(n_tiles, tid * per), end = min(n_tiles, begin + per);
  uint32_t c[kPer];
  uint32_t sum = 0;
#pragma unroll
  for (uint32_t j = 0; j < kPer; ++j) {
    c[j] = begin + j < end ? row[begin + j] : 0u;
    sum += c[j];
  }
  uint32_t incl = sum;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= uint32_t(o))
      incl += u;
  }
  __syncthreads(); // warp_sums is reused
  if (lane == 31)
    warp_sums[warp] = incl;
  __syncthreads();
  uint32_t run = base + incl - sum;
  for (uint32_t w = 0; w < warp; ++w)
    run += warp_sums[w];
#pragma unroll
  for (uint32_t j = 0; j < kPer; ++j)
    if (begin + j < end) {
      row[begin + j] = run;
      run += c[j];
    }
}

__global__ void __launch_bounds__(32 * kWarpsPerCta)
radix_scatter_kernel(const unsigned long long* __restrict__ keys_in, const uint32_t* __restrict__ idx_in,
                     unsigned long long* __restrict__ keys_out, uint32_t* __restrict__ idx_out, uint32_t n, uint32_t shift,
                     uint32_t tile_elems, uint32_t n_tiles, const uint32_t* __restrict__ hist)
{
  __shared__ uint32_t base[kWarpsPerCta][kBuckets];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tile = blockIdx.x * kWarpsPerCta + warp;
  if (tile >= n_tiles)
    return;
  for (int d = lane; d < kBuckets; d += 32)
    base[warp][d] = hist[size_t(d) * n_tiles + tile];
  __syncwarp();
  const uint32_t begin = tile * tile_elems, end = min(n, begin + tile_elems);
  const uint32_t lt = (1u << lane) - 1u;
  for (uint32_t i0 = begin; i0 < end; i0 += 32) { // rounds of 32 consecutive records: order inside the tile is kept
    const uint32_t i = i0 + lane;
    const bool active = i < end;
    const uint32_t mask = __ballot_sync(0xFFFFFFFFu, active);
    if (active) {
      const unsigned long long key = keys_in[i];
      const uint32_t val = idx_in[i];
      const uint32_t d = digit_of(key, shift);
      const uint32_t peers = __match_any_sync(mask, d);
      const uint32_t pos = base[warp][d] + __popc(peers & lt);
      __syncwarp(mask);
      if ((peers & lt) == 0) // first lane of its digit in this round
        base[warp][d] += __popc(peers);
      keys_out[pos] = key;
      idx_out[pos] = val;
    }
    __syncwarp();
  }
}

__global__ void
tp_gather_kernel(const swtpg_tp* __restrict__ tps, const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ idx, uint32_t n,
                 swtpg_tp* __restrict__ sorted, uint32_t* __restrict__ tie_flag)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const uint4* src = reinterpret_cast<const uint4*>(tps + idx[i]);
  const uint4 a = src[0], b = src[1];
  uint4* dst = reinterpret_cast<uint4*>(sorted + i);
  dst[0] = a;
  dst[1] = b;
  if (i && keys[i] == keys[i - 1])
    *tie_flag = 1u;
}

unsigned
bit_length(unsigned long long v)
{
  unsigned b = 0;
  while (v) {
    ++b;
    v >>= 1;
  }
  return b;
}

} // namespace

struct TpSorter
{
  std::mutex mu;
  size_t cap = 0;
  unsigned long long* keys[2] = { nullptr, nullptr };
  uint32_t* idx[2] = { nullptr, nullptr };
  swtpg_tp* sorted = nullptr;
  uint32_t* hist = nullptr;          // [kBuckets][kMaxTiles], followed by the digit totals of up to 8 passes: [8][kBuckets]
  unsigned long long* d_scalars = nullptr; // {min, max, tie flag, OR of the bits in which time_start values differ}
  unsigned long long* h_scalars = nullptr; // pinned
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::atomic<uint64_t> last_us{ 0 }, total_us{ 0 }, calls{ 0 }, host_fallbacks{ 0 };

  void release()
  {
    for (int i = 0; i < 2; ++i) {
      if (keys[i]) cudaFree(keys[i]);
      if (idx[i]) cudaFree(idx[i]);
      keys[i] = nullptr;
      idx[i] = nullptr;
    }
    if (sorted) cudaFree(sorted);
    sorted = nullptr;
    cap = 0;
  }
  ~TpSorter()
  {
    release();
    if (hist) cudaFree(hist);
    if (d_scalars) cudaFree(d_scalars);
    if (h_scalars) cudaFreeHost(h_scalars);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
  }
};

TpSorter*
sorter_create()
{
  return new TpSorter;
}

void
sorter_destroy(TpSorter* s)
{
  delete s;
}

void
sorter_stats(const TpSorter* s, double* last_ms, double* total_ms, uint64_t* calls, uint64_t* host_fallbacks)
{
  if (last_ms) *last_ms = s ? double(s->last_us.load()) * 1e-3 : -1.0;
  if (total_ms) *total_ms = s ? double(s->total_us.load()) * 1e-3 : 0.0;
  if (calls) *calls = s ? s->calls.load() : 0;
  if (host_fallbacks) *host_fallbacks = s ? s->host_fallbacks.load() : 0;
}

// Orders d_tps[0..n) on stream `s` (which must already be ordered after the kernel that produced the list). On success *out is
// the ordered list on the device (the sorter's buffer, valid until the next call: the caller holds `lock` until its copy of the
// list has been enqueued AND completed) and *finish_on_host says whether the host tie-break / full host sort still has to run on
// the copy (equal keys seen, or keys wider than 64 bits, in which case *out is the unordered list itself).
swtpg_status
sort_tps_device(swtpg_handle* h, TpSorter* st, const swtpg_tp* d_tps, size_t n, cudaStream_t s, const swtpg_tp** out, bool* finish_on_host,
                std::unique_lock<std::mutex>* lock)
{
  *out = d_tps;
  *finish_on_host = false;
  if (n < 2)
    return SWTPG_OK;
  if (n > 0xFFFFFFF0ull) {
    *finish_on_host = true;
    return SWTPG_OK;
  }
  *lock = std::unique_lock<std::mutex>(st->mu);
  if (!st->hist) {
    SW_CUDA(h, cudaMalloc(&st->hist, (size_t(kBuckets) * kMaxTiles + 8 * kBuckets) * sizeof(uint32_t)));
    SW_CUDA(h, cudaMalloc(&st->d_scalars, 4 * sizeof(unsigned long long)));
    SW_CUDA(h, cudaMallocHost(&st->h_scalars, 4 * sizeof(unsigned long long)));
    SW_CUDA(h, cudaEventCreate(&st->ev0));
    SW_CUDA(h, cudaEventCreate(&st->ev1));
  }
  if (st->cap < n) { // grow-only scratch: 56 bytes per record
    SW_CUDA(h, cudaStreamSynchronize(s));
    st->release();
    const size_t cap = std::min<size_t>(std::max<size_t>(n + n / 4, 1u << 16), std::max<size_t>(h->tp_capacity, n));
    for (int i = 0; i < 2; ++i) {
      SW_CUDA(h, cudaMalloc(&st->keys[i], cap * sizeof(unsigned long long)));
      SW_CUDA(h, cudaMalloc(&st->idx[i], cap * sizeof(uint32_t)));
    }
    SW_CUDA(h, cudaMalloc(&st->sorted, cap * sizeof(swtpg_tp)));
    st->cap = cap;
  }
  const uint32_t N = uint32_t(n);
  SW_CUDA(h, cudaEventRecord(st->ev0, s));
  SW_CUDA(h, cudaMemsetAsync(st->d_scalars, 0xFF, sizeof(unsigned long long), s));
  SW_CUDA(h, cudaMemsetAsync(st->d_scalars + 1, 0, 3 * sizeof(unsigned long long), s));
  uint32_t* totals = st->hist + size_t(kBuckets) * kMaxTiles;
  SW_CUDA(h, cudaMemsetAsync(totals, 0, 8 * kBuckets * sizeof(uint32_t), s));
  tp_minmax_kernel<<<std::min<uint32_t>((N + 255) / 256, 592), 256, 0, s>>>(d_tps, N, st->d_scalars);
  SW_CUDA(h, cudaGetLastError());
  SW_CUDA(h, cudaMemcpyAsync(st->h_scalars, st->d_scalars, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  SW_CUDA(h, cudaStreamSynchronize(s));
  const unsigned long long tmin = st->h_scalars[0], tmax = st->h_scalars[1], differ = st->h_scalars[3];
  unsigned tz = 0; // low bits shared by every time_start
  while (tz < 63 && differ && !((differ >> tz) & 1u))
    ++tz;
  const unsigned cb = bit_length(h->channels - 1), lb = bit_length(h->cfg.n_links - 1), tb = bit_length((tmax - tmin) >> tz), bits = cb + lb + tb;
  if (bits > 64) {
    st->host_fallbacks++;
    *finish_on_host = true;
    return SWTPG_OK;
  }
  tp_keys_kernel<<<(N + 255) / 256, 256, 0, s>>>(d_tps, N, tmin, tz, lb, cb, st->keys[0], st->idx[0]);
  SW_CUDA(h, cudaGetLastError());
  uint32_t tile = std::max<uint32_t>(512, (N + kMaxTiles - 1) / kMaxTiles);
  tile = (tile + 31) & ~31u;
  const uint32_t n_tiles = (N + tile - 1) / tile, ctas = (n_tiles + kWarpsPerCta - 1) / kWarpsPerCta;
  int cur = 0;
  for (unsigned shift = 0; shift < bits; shift += kDigitBits, totals += kBuckets) {
    radix_hist_kernel<<<ctas, 32 * kWarpsPerCta, 0, s>>>(st->keys[cur], N, shift, tile, n_tiles, st->hist, totals);
    radix_scan_kernel<<<kBuckets, 256, 0, s>>>(st->hist, totals, n_tiles);
    radix_scatter_kernel<<<ctas, 32 * kWarpsPerCta, 0, s>>>(st->keys[cur], st->idx[cur], st->keys[cur ^ 1], st->idx[cur ^ 1], N, shift, tile, n_tiles,
                                                            st->hist);
    SW_CUDA(h, cudaGetLastError());
    cur ^= 1;
  }
  tp_gather_kernel<<<(N + 255) / 256, 256, 0, s>>>(d_tps, st->keys[cur], st->idx[cur], N, st->sorted,
                                                    reinterpret_cast<uint32_t*>(st->d_scalars + 2));
  SW_CUDA(h, cudaGetLastError());
  SW_CUDA(h, cudaMemcpyAsync(st->h_scalars + 2, st->d_scalars + 2, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  SW_CUDA(h, cudaEventRecord(st->ev1, s));
  SW_CUDA(h, cudaStreamSynchronize(s));
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, st->ev0, st->ev1) == cudaSuccess) {
    st->last_us = uint64_t(ms * 1000.f);
    st->total_us += uint64_t(ms * 1000.f);
  }
  st->calls++;
  *finish_on_host = st->h_scalars[2] != 0;
  if (*finish_on_host)
    st->host_fallbacks++;
  *out = st->sorted;
  return SWTPG_OK;
}

} // namespace swtpg_internal
