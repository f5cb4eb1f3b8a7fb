// Host runtime behind include/swtpg.h: device buffers, kernel selection, the batch entry points and the pinned
// multi-stream staging ring of the streaming path. Plain CUDA runtime; no torch, no CPU compute fallback.
#include "../../include/swtpg.h"
#include "swtpg_kernels.cuh"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

using namespace swtpg;

// Layouts the ctypes binding (fdreadoutlibs_b200/_lib.py, frames.py) relies on.
static_assert(sizeof(swtpg_tp) == 32, "swtpg_tp must stay 32 bytes (two 16-byte device stores)");
static_assert(sizeof(swtpg_config) == 68, "swtpg_config layout changed: bump SWTPG_ABI_VERSION and the bindings");
static_assert(sizeof(swtpg_channel_state) == 48, "swtpg_channel_state layout changed");
static_assert(sizeof(swtpg_counters) == 64, "swtpg_counters layout changed");

namespace {

thread_local std::string g_create_error;

enum SlotState : int
{
  kFilling = 0,
  kCopying = 1,  // H2D + kernel + count D2H enqueued
  kFetching = 2, // TP D2H enqueued
  kReady = 3     // TPs in h_tps, waiting for poll
};

struct Slot
{
  uint8_t* h_frames = nullptr; // pinned [n_links][max_units][unit_bytes]
  uint8_t* d_frames = nullptr;
  swtpg_tp* d_tps = nullptr;
  swtpg_tp* h_tps = nullptr;   // pinned
  unsigned* d_count = nullptr;
  unsigned* h_count = nullptr; // pinned
  uint32_t* h_nunits = nullptr; // pinned [n_links]
  uint32_t* d_nunits = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_h2d = nullptr, ev_kernel = nullptr, ev_count = nullptr, ev_tps = nullptr;
  std::atomic<int> state{ kFilling };
  std::atomic<uint64_t> batch{ 0 };      // batch index this slot currently holds
  std::atomic<uint32_t> remaining{ 0 };  // units still missing before auto-dispatch
  uint32_t n_ready = 0, n_taken = 0;
  // zero-copy ingest (swtpg_register_buffer): where each unit of the batch lies if it was NOT copied into h_frames
  std::vector<const uint8_t*> borrowed;    // [n_links][max_units], nullptr = staged copy
  std::atomic<uint32_t> n_borrowed{ 0 };
};

} // namespace

struct swtpg_handle
{
  swtpg_config cfg{};
  uint32_t unit_bytes = 0, channels = 0, ticks = 0, groups_per_link = 0, n_groups = 0;
  uint32_t tp_capacity = 0;
  bool fast_simple = false, fast_fir = false, fast_rs = false, fast_rs_wib2 = false, fast_fir_any = false;
  bool started = false;

  cudaStream_t stream = nullptr; // batch path + all kernels (state is carried batch to batch: kernels are ordered)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;
  cudaStream_t last_stream = nullptr;

  uint32_t* d_state = nullptr;
  uint32_t* d_flags = nullptr;
  uint32_t* d_link_cursor = nullptr; // {claimed, finished}: dynamic link hand-out of the WIBEth kernel, self-resetting
  swtpg_tp* d_tps = nullptr;
  unsigned* d_count = nullptr;
  unsigned* h_count = nullptr;
  uint32_t* d_nunits = nullptr;
  uint32_t* h_nunits = nullptr;
  uint8_t* d_frames = nullptr;
  size_t d_frames_bytes = 0;
  int16_t* d_ped = nullptr;
  int16_t* d_wav = nullptr;
  size_t d_dump_elems = 0;
  uint16_t* h_rs_factor = nullptr; // [n_links][channels] or null

  // streaming path
  std::vector<std::unique_ptr<Slot>> slots;
  std::atomic<bool> slots_ready{ false };
  std::unique_ptr<std::atomic<uint64_t>[]> submitted; // units per link
  std::mutex dispatch_mu;
  uint64_t next_dispatch = 0; // next batch index to dispatch (batches complete in order)
  uint64_t next_poll = 0;     // next batch index to hand to poll

  // host ranges registered for zero-copy ingest; link_range caches the last hit per link (touched by that link's thread only)
  struct HostRange { uintptr_t lo = 0, hi = 0; };
  std::vector<HostRange> ranges;
  std::mutex ranges_mu;
  std::unique_ptr<HostRange[]> link_range;
  std::atomic<uint64_t> ranges_epoch{ 0 };
  std::unique_ptr<uint64_t[]> link_range_epoch;

  // bounce pipeline of swtpg_process_host for pageable sources: per worker two pinned buffers, a stream and events
  struct Bounce
  {
    uint8_t* buf[2] = { nullptr, nullptr };
    cudaEvent_t free_ev[2] = { nullptr, nullptr };
    cudaEvent_t done = nullptr;
    cudaStream_t stream = nullptr;
  };
  std::vector<Bounce> bounce;

  swtpg_counters counters{};
  std::atomic<uint64_t> submit_busy{ 0 };
  mutable std::string last_error;
};

namespace {

#define SW_CUDA(h, call)                                                                                                          \
  do {                                                                                                                            \
    cudaError_t e_ = (call);                                                                                                      \
    if (e_ != cudaSuccess) {                                                                                                      \
      char buf_[512];                                                                                                             \
      snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);                    \
      if (h)                                                                                                                      \
        (h)->last_error = buf_;                                                                                                   \
      else                                                                                                                        \
        g_create_error = buf_;                                                                                                    \
      return SWTPG_ERR_CUDA;                                                                                                      \
    }                                                                                                                             \
  } while (0)

swtpg_status
fail(swtpg_handle* h, swtpg_status s, const char* msg)
{
  if (h)
    h->last_error = msg;
  else
    g_create_error = msg;
  return s;
}

// ---- kernel launch table ------------------------------------------------------------------------------------------
// Geometry of the WIBEth kernel: WARPS links per CTA, per-warp ring of NSTAGE stages of CHUNK ticks (112 B each).
template<int WARPS, int NSTAGE, int CHUNK, int MIN_CTAS = 1>
struct Geo
{
  static constexpr int warps = WARPS, stages = NSTAGE, chunk = CHUNK, min_ctas = MIN_CTAS;
  static constexpr size_t smem = WibEthSmem<WARPS, NSTAGE, CHUNK>::total;
};
// Default: one link-warp per CTA with a ring of 2 stages x 32 ticks (7 KB) + 2.5 KB hit staging. Measured best of the
// geometries below on B200 (profiles/r01_geometry_sweep.txt): the kernel is issue-bound, so deeper rings do not help, and
// single-warp CTAs let the block scheduler spread the 20 resident warps per SM evenly over the four sub-partitions.
using GeoDefault = Geo<1, 2, 32>;

template<class Algo, bool DUMP, class G>
cudaError_t
launch_wibeth_geo(const KernelParams& kp, cudaStream_t s)
{
  auto k = wibeth_kernel<Algo, G::warps, G::stages, G::chunk, DUMP, G::min_ctas>;
  // Persistent grid: as many CTAs as the device holds at once (SMs x resident CTAs per SM); warps walk the links.
  static int resident[64]; // per instantiation and device: CTAs the whole GPU can hold, 0 = not queried yet
  static int sm_count[64];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess)
    return e;
  if (dev < 0 || dev >= 64)
    return cudaErrorInvalidDevice;
  if (resident[dev] == 0) {
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(G::smem));
    if (e != cudaSuccess)
      return e;
    int per_sm = 0, sms = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, G::warps * 32, G::smem);
    if (e != cudaSuccess)
      return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess)
      return e;
    if (const char* cap = getenv("SWTPG_CTAS_PER_SM")) // tuning aid
      per_sm = std::min(per_sm, std::max(1, atoi(cap)));
    sm_count[dev] = std::max(1, sms);
    resident[dev] = std::max(1, per_sm * sms);
  }
  // Persistent warps claim links dynamically (wibeth_kernel), so the grid only has to load every SM sub-partition alike:
  // a multiple of 4 single-warp CTAs per SM (20 on B200: 5 warps per sub-partition; a 21st warp would make one
  // sub-partition of every SM 20 % slower than its neighbours — profiles/README.md).
  const unsigned per_sm = unsigned(resident[dev]) / unsigned(sm_count[dev]);
  unsigned even = G::warps == 1 && per_sm >= 4 ? per_sm / 4 * 4 : per_sm;
  if (G::warps == 1 && Algo::kWarpsPerSm > 0) // the policy's measured optimum, if the device holds that many
    even = std::min<unsigned>(even, unsigned(Algo::kWarpsPerSm));
  unsigned warps = std::min<unsigned>(kp.n_links, even * unsigned(sm_count[dev]) * G::warps);
  static const int warps_override = [] { const char* e = getenv("SWTPG_WARPS"); return e ? atoi(e) : 0; }(); // tuning aid
  if (warps_override > 0)
    warps = std::min<unsigned>(unsigned(warps_override), kp.n_links);
  const unsigned grid = std::min<unsigned>((warps + G::warps - 1) / G::warps, unsigned(resident[dev]));
  k<<<grid, G::warps * 32, G::smem, s>>>(kp);
  return cudaGetLastError();
}

// CTA form of the WIBEth kernel (wibeth_quad_kernel): 4 consumer warps + 1 producer warp per quad of links.
template<class Algo, bool DUMP>
cudaError_t
launch_wibeth_quad(const KernelParams& kp, cudaStream_t s)
{
#ifndef SWTPG_QUAD_STAGES
#define SWTPG_QUAD_STAGES 2
#define SWTPG_QUAD_CHUNK 32
#endif
  constexpr int kStages = SWTPG_QUAD_STAGES, kChunk = SWTPG_QUAD_CHUNK;
  auto k = wibeth_quad_kernel<Algo, kStages, kChunk, DUMP>;
  constexpr size_t smem = WibEthQuadSmem<kStages, kChunk>::total;
  static int resident[64];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess)
    return e;
  if (dev < 0 || dev >= 64)
    return cudaErrorInvalidDevice;
  if (resident[dev] == 0) {
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess)
      return e;
    int per_sm = 0, sms = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, (kQuad + 1) * 32, smem);
    if (e != cudaSuccess)
      return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess)
      return e;
    if (const char* cap = getenv("SWTPG_CTAS_PER_SM")) // tuning aid
      per_sm = std::min(per_sm, std::max(1, atoi(cap)));
    else if (Algo::kQuadCtasPerSm > 0) // the policy's measured optimum, if the device holds that many
      per_sm = std::min(per_sm, Algo::kQuadCtasPerSm);
    resident[dev] = std::max(1, per_sm * sms);
  }
  const unsigned quads = (kp.n_links + kQuad - 1) / kQuad;
  const unsigned grid = std::min<unsigned>(quads, unsigned(resident[dev])); // persistent CTAs claim quads dynamically
  k<<<grid, (kQuad + 1) * 32, smem, s>>>(kp);
  return cudaGetLastError();
}

// Which form of the WIBEth kernel a policy runs is a measured choice (profiles/r01_quad_vs_warp.txt): the CTA form is 14 % faster
// for FIR + IQR (ptxas needs 64 registers instead of 109 for it, so 20 consumer warps fit an SM), the one-warp-per-CTA form
// 2-5 % faster for SimpleThreshold and the running sums (the quad's lock-step costs more than the producer bookkeeping it
// saves). SWTPG_WIBETH_KERNEL=warp forces the latter.
template<class Algo, bool DUMP>
cudaError_t
launch_wibeth(const KernelParams& kp, cudaStream_t s)
{
  if constexpr (Algo::kQuadCtasPerSm > 0) {
    static const bool warp_form = [] {
      const char* e = getenv("SWTPG_WIBETH_KERNEL");
      return e && std::string(e) == "warp";
    }();
    if (!warp_form)
      return launch_wibeth_quad<Algo, DUMP>(kp, s);
  }
  return launch_wibeth_geo<Algo, DUMP, GeoDefault>(kp, s);
}

// WIB2: one CTA (4 consumer warps + 1 producer warp) per link at a time, ring of 4 superchunks; persistent over links.
template<class Algo, bool DUMP>
cudaError_t
launch_wib2(const KernelParams& kp, cudaStream_t s)
{
#ifndef SWTPG_WIB2_STAGES
#define SWTPG_WIB2_STAGES 4
#endif
  constexpr int kStages = SWTPG_WIB2_STAGES;
  auto k = wib2_kernel<Algo, kStages, DUMP>;
  constexpr size_t smem = Wib2Smem<kStages>::total;
  static int resident[64];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess)
    return e;
  if (dev < 0 || dev >= 64)
    return cudaErrorInvalidDevice;
  if (resident[dev] == 0) {
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess)
      return e;
    int per_sm = 0, sms = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, (kWib2Warps + 1) * 32, smem);
    if (e != cudaSuccess)
      return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess)
      return e;
    resident[dev] = std::max(1, per_sm * sms);
  }
  const unsigned grid = std::min<unsigned>(kp.n_links, unsigned(resident[dev])); // persistent CTAs claim links dynamically
  k<<<grid, (kWib2Warps + 1) * 32, smem, s>>>(kp); // 4 consumer warps + the producer warp
  return cudaGetLastError();
}

template<bool DUMP>
cudaError_t
launch(const swtpg_handle* h, const KernelParams& kp, cudaStream_t s)
{
  if (h->cfg.format == SWTPG_FORMAT_WIB2) {
    switch (h->cfg.algorithm) {
      case SWTPG_ALGO_SIMPLE_THRESHOLD:
        return h->fast_simple ? launch_wib2<PackedSimpleWib2, DUMP>(kp, s)
                              : launch_wib2<ScalarAlgo<SWTPG_ALGO_SIMPLE_THRESHOLD, true>, DUMP>(kp, s);
      case SWTPG_ALGO_FIR_IQR:
        return h->fast_fir       ? launch_wib2<PackedFirIqr, DUMP>(kp, s)
               : h->fast_fir_any ? launch_wib2<PackedFirIqrAnyTaps, DUMP>(kp, s)
                                 : launch_wib2<ScalarAlgo<SWTPG_ALGO_FIR_IQR, true>, DUMP>(kp, s);
      case SWTPG_ALGO_ABS_RS:
        return h->fast_rs_wib2 ? launch_wib2<PackedRsIqrWib2, DUMP>(kp, s) : launch_wib2<ScalarAlgo<SWTPG_ALGO_ABS_RS, true>, DUMP>(kp, s);
      default: return cudaErrorNotSupported;
    }
  }
  if (h->cfg.format == SWTPG_FORMAT_WIBETH) {
    switch (h->cfg.algorithm) {
      case SWTPG_ALGO_SIMPLE_THRESHOLD:
        return h->fast_simple ? launch_wibeth<PackedSimpleWibEth, DUMP>(kp, s)
                              : launch_wibeth<ScalarAlgo<SWTPG_ALGO_SIMPLE_THRESHOLD, false>, DUMP>(kp, s);
      case SWTPG_ALGO_ABS_RS:
        return h->fast_rs ? launch_wibeth<PackedRsWibEth<false>, DUMP>(kp, s) : launch_wibeth<ScalarAlgo<SWTPG_ALGO_ABS_RS, false>, DUMP>(kp, s);
      case SWTPG_ALGO_STANDARD_RS:
        return h->fast_rs ? launch_wibeth<PackedRsWibEth<true>, DUMP>(kp, s)
                          : launch_wibeth<ScalarAlgo<SWTPG_ALGO_STANDARD_RS, false>, DUMP>(kp, s);
      case SWTPG_ALGO_FIR_IQR:
        return h->fast_fir       ? launch_wibeth<PackedFirIqr, DUMP>(kp, s)
               : h->fast_fir_any ? launch_wibeth<PackedFirIqrAnyTaps, DUMP>(kp, s)
                                 : launch_wibeth<ScalarAlgo<SWTPG_ALGO_FIR_IQR, false>, DUMP>(kp, s);
    }
  }
  return cudaErrorNotSupported;
}

KernelParams
make_params(const swtpg_handle* h, const void* d_frames, const uint32_t* d_nunits, uint32_t stride, swtpg_tp* d_tps, unsigned* d_count,
            int16_t* ped, int16_t* wav)
{
  KernelParams kp{};
  kp.frames = static_cast<const uint8_t*>(d_frames);
  kp.n_units = d_nunits;
  kp.units_stride = stride;
  kp.n_links = h->cfg.n_links;
  kp.state = h->d_state;
  kp.group_flags = h->d_flags;
  kp.link_cursor = h->d_link_cursor;
  kp.sink.buf = d_tps;
  kp.sink.count = d_count;
  kp.sink.cap = h->tp_capacity;
  kp.pedestal_out = ped;
  kp.waveform_out = wav;
  kp.threshold = h->cfg.threshold;
  kp.acc_limit = h->cfg.frugal_acc_limit;
  kp.rs_scale = int16_t(h->cfg.rs_scale_factor);
  kp.tap_exponent = h->cfg.tap_exponent;
  for (int i = 0; i < 8; ++i)
    kp.taps[i] = h->cfg.fir_taps[i];
  kp.wib2_adc_offset = h->cfg.wib2_adc_offset;
  static const bool force_exact = [] { const char* e = getenv("SWTPG_FIR_FORCE_EXACT"); return e && atoi(e) != 0; }();
  kp.debug_flags = force_exact ? 1u : 0u;
  return kp;
}

swtpg_status
reset_state(swtpg_handle* h)
{
  // Fresh zeroed ChanState (wibeth/tpg/ProcessingInfo.hpp:23-40) + per-channel RS memory factor (setState :131)
  std::vector<uint32_t> st(size_t(h->n_groups) * kStateWordsPerGroup, 0u);
  for (uint32_t g = 0; g < h->n_groups; ++g) {
    const uint32_t link = g / h->groups_per_link, sub = g % h->groups_per_link;
    for (uint32_t lane = 0; lane < 32; ++lane) {
      uint32_t lo = h->cfg.rs_memory_factor, hi = h->cfg.rs_memory_factor;
      if (h->h_rs_factor) {
        const uint16_t* f = h->h_rs_factor + size_t(link) * h->channels + sub * 64 + 2 * lane;
        lo = f[0];
        hi = f[1];
      }
      st[size_t(g) * kStateWordsPerGroup + SV_RS_FACTOR * 32 + lane] = lo | (hi << 16);
    }
  }
  SW_CUDA(h, cudaMemcpyAsync(h->d_state, st.data(), st.size() * 4, cudaMemcpyHostToDevice, h->stream));
  SW_CUDA(h, cudaMemsetAsync(h->d_flags, 0, size_t(h->n_groups) * 4, h->stream));
  SW_CUDA(h, cudaMemsetAsync(h->d_link_cursor, 0, 2 * sizeof(uint32_t), h->stream));
  SW_CUDA(h, cudaStreamSynchronize(h->stream));
  return SWTPG_OK;
}

swtpg_status
check_batch_args(swtpg_handle* h, const void* frames, const uint32_t* n_units, uint32_t stride)
{
  if (!h)
    return SWTPG_ERR_INVALID_ARG;
  if (!h->started)
    return fail(h, SWTPG_ERR_STATE, "swtpg_start has not been called");
  if (!frames && stride)
    return fail(h, SWTPG_ERR_INVALID_ARG, "frames is NULL");
  if (stride > h->cfg.max_units)
    return fail(h, SWTPG_ERR_INVALID_ARG, "units_stride exceeds cfg.max_units");
  if (n_units)
    for (uint32_t l = 0; l < h->cfg.n_links; ++l)
      if (n_units[l] > stride)
        return fail(h, SWTPG_ERR_INVALID_ARG, "n_units[link] exceeds units_stride");
  return SWTPG_OK;
}

// Enqueue one batch on `s`: n_units upload, counter reset, kernel (timed with events on the same stream).
swtpg_status
enqueue_batch(swtpg_handle* h, const void* d_frames, const uint32_t* n_units, uint32_t stride, cudaStream_t s, bool dump)
{
  const uint32_t* d_nu = nullptr;
  uint64_t units = 0;
  if (n_units) {
    memcpy(h->h_nunits, n_units, size_t(h->cfg.n_links) * 4);
    SW_CUDA(h, cudaMemcpyAsync(h->d_nunits, h->h_nunits, size_t(h->cfg.n_links) * 4, cudaMemcpyHostToDevice, s));
    d_nu = h->d_nunits;
    for (uint32_t l = 0; l < h->cfg.n_links; ++l)
      units += n_units[l];
  } else {
    units = uint64_t(stride) * h->cfg.n_links;
  }
  SW_CUDA(h, cudaMemsetAsync(h->d_count, 0, sizeof(unsigned), s));
  if (stride) {
    KernelParams kp = make_params(h, d_frames, d_nu, stride, h->d_tps, h->d_count, dump ? h->d_ped : nullptr, dump ? h->d_wav : nullptr);
    SW_CUDA(h, cudaEventRecord(h->ev0, s));
    SW_CUDA(h, dump ? launch<true>(h, kp, s) : launch<false>(h, kp, s));
    SW_CUDA(h, cudaEventRecord(h->ev1, s));
    h->timed = true;
  }
  h->last_stream = s;
  h->counters.units_processed += units;
  h->counters.samples_processed += units * h->channels * h->ticks;
  h->counters.batches++;
  return SWTPG_OK;
}

swtpg_status
fetch(swtpg_handle* h, cudaStream_t s, swtpg_tp* out, size_t cap, size_t* n_out)
{
  SW_CUDA(h, cudaMemcpyAsync(h->h_count, h->d_count, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
  SW_CUDA(h, cudaStreamSynchronize(s));
  const size_t found = *h->h_count;
  const size_t stored = std::min<size_t>(found, h->tp_capacity);
  const size_t n = std::min(stored, cap);
  if (n && out)
    SW_CUDA(h, cudaMemcpyAsync(out, h->d_tps, n * sizeof(swtpg_tp), cudaMemcpyDeviceToHost, s));
  SW_CUDA(h, cudaStreamSynchronize(s));
  if (n_out)
    *n_out = found;
  h->counters.tps_emitted += found;
  h->counters.d2h_bytes += n * sizeof(swtpg_tp) + sizeof(unsigned);
  if (found > stored)
    h->counters.tps_dropped_overflow += found - stored;
  if (found > n)
    return fail(h, SWTPG_ERR_OVERFLOW, "more TPs than capacity");
  return SWTPG_OK;
}

void
free_slot(Slot& s)
{
  if (s.h_frames) cudaFreeHost(s.h_frames);
  if (s.d_frames) cudaFree(s.d_frames);
  if (s.d_tps) cudaFree(s.d_tps);
  if (s.h_tps) cudaFreeHost(s.h_tps);
  if (s.d_count) cudaFree(s.d_count);
  if (s.h_count) cudaFreeHost(s.h_count);
  if (s.h_nunits) cudaFreeHost(s.h_nunits);
  if (s.d_nunits) cudaFree(s.d_nunits);
  if (s.ev_h2d) cudaEventDestroy(s.ev_h2d);
  if (s.ev_kernel) cudaEventDestroy(s.ev_kernel);
  if (s.ev_count) cudaEventDestroy(s.ev_count);
  if (s.ev_tps) cudaEventDestroy(s.ev_tps);
  if (s.stream) cudaStreamDestroy(s.stream);
}

// Streaming ring is allocated on first swtpg_submit (the batch entry points never need it).
swtpg_status
ensure_slots(swtpg_handle* h)
{
  if (h->slots_ready.load(std::memory_order_acquire))
    return SWTPG_OK;
  const uint32_t n = h->cfg.n_slots;
  const size_t fbytes = size_t(h->cfg.n_links) * h->cfg.max_units * h->unit_bytes;
  std::vector<std::unique_ptr<Slot>> slots;
  for (uint32_t i = 0; i < n; ++i) {
    auto s = std::make_unique<Slot>();
    SW_CUDA(h, cudaMallocHost(&s->h_frames, fbytes));
    SW_CUDA(h, cudaMalloc(&s->d_frames, fbytes));
    SW_CUDA(h, cudaMalloc(&s->d_tps, size_t(h->tp_capacity) * sizeof(swtpg_tp)));
    SW_CUDA(h, cudaMallocHost(&s->h_tps, size_t(h->tp_capacity) * sizeof(swtpg_tp)));
    SW_CUDA(h, cudaMalloc(&s->d_count, sizeof(unsigned)));
    SW_CUDA(h, cudaMallocHost(&s->h_count, sizeof(unsigned)));
    SW_CUDA(h, cudaMallocHost(&s->h_nunits, size_t(h->cfg.n_links) * 4));
    SW_CUDA(h, cudaMalloc(&s->d_nunits, size_t(h->cfg.n_links) * 4));
    SW_CUDA(h, cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    SW_CUDA(h, cudaEventCreateWithFlags(&s->ev_h2d, cudaEventDisableTiming));
    SW_CUDA(h, cudaEventCreateWithFlags(&s->ev_kernel, cudaEventDisableTiming));
    SW_CUDA(h, cudaEventCreateWithFlags(&s->ev_count, cudaEventDisableTiming));
    SW_CUDA(h, cudaEventCreateWithFlags(&s->ev_tps, cudaEventDisableTiming));
    s->borrowed.assign(size_t(h->cfg.n_links) * h->cfg.max_units, nullptr);
    s->batch.store(i);
    s->remaining.store(h->cfg.n_links * h->cfg.max_units);
    s->state.store(kFilling);
    slots.push_back(std::move(s));
  }
  h->slots = std::move(slots);
  h->slots_ready.store(true, std::memory_order_release);
  return SWTPG_OK;
}

// Dispatch batch `b` (its slot is full or being flushed). Caller holds dispatch_mu. n_units == nullptr: full batch.
swtpg_status
dispatch_slot(swtpg_handle* h, Slot& s, const uint32_t* n_units)
{
  const uint32_t stride = h->cfg.max_units;
  const size_t fbytes = size_t(h->cfg.n_links) * stride * h->unit_bytes;
  const uint32_t* d_nu = nullptr;
  uint64_t units = uint64_t(stride) * h->cfg.n_links;
  // H2D on the slot's own stream (overlaps the previous batch's kernel), kernel on the handle's compute stream
  // (state is carried: kernels must run in batch order), TP count + TPs back on the slot's stream.
  const bool any_borrowed = s.n_borrowed.load(std::memory_order_acquire) != 0;
  if (n_units) {
    units = 0;
    for (uint32_t l = 0; l < h->cfg.n_links; ++l) {
      s.h_nunits[l] = n_units[l];
      units += n_units[l];
    }
    SW_CUDA(h, cudaMemcpyAsync(s.d_nunits, s.h_nunits, size_t(h->cfg.n_links) * 4, cudaMemcpyHostToDevice, s.stream));
    d_nu = s.d_nunits;
  }
  if (any_borrowed) {
    // Zero-copy ingest: the copy engine reads borrowed units where they lie (registered host memory), one async copy per
    // contiguous run — a link's superchunk is one run unless the latency buffer wrapped inside it; staged units come from
    // the slot's pinned buffer as before.
    for (uint32_t l = 0; l < h->cfg.n_links; ++l) {
      const uint32_t nu = n_units ? n_units[l] : stride;
      const size_t row = size_t(l) * stride;
      uint32_t u = 0;
      while (u < nu) {
        const uint8_t* src = s.borrowed[row + u];
        const bool staged = src == nullptr;
        if (staged)
          src = s.h_frames + (row + u) * h->unit_bytes;
        uint32_t v = u + 1;
        while (v < nu) {
          const uint8_t* nxt = s.borrowed[row + v];
          if (staged ? nxt != nullptr : nxt != src + size_t(v - u) * h->unit_bytes)
            break;
          ++v;
        }
        SW_CUDA(h, cudaMemcpyAsync(s.d_frames + (row + u) * h->unit_bytes, src, size_t(v - u) * h->unit_bytes, cudaMemcpyHostToDevice,
                                   s.stream));
        u = v;
      }
    }
    h->counters.h2d_bytes += units * h->unit_bytes;
  } else if (n_units) {
    // ragged: copy each link's valid prefix only
    for (uint32_t l = 0; l < h->cfg.n_links; ++l)
      if (n_units[l])
        SW_CUDA(h, cudaMemcpyAsync(s.d_frames + size_t(l) * stride * h->unit_bytes, s.h_frames + size_t(l) * stride * h->unit_bytes,
                                   size_t(n_units[l]) * h->unit_bytes, cudaMemcpyHostToDevice, s.stream));
    h->counters.h2d_bytes += units * h->unit_bytes;
  } else {
    SW_CUDA(h, cudaMemcpyAsync(s.d_frames, s.h_frames, fbytes, cudaMemcpyHostToDevice, s.stream));
    h->counters.h2d_bytes += fbytes;
  }
  SW_CUDA(h, cudaMemsetAsync(s.d_count, 0, sizeof(unsigned), s.stream));
  SW_CUDA(h, cudaEventRecord(s.ev_h2d, s.stream));
  SW_CUDA(h, cudaStreamWaitEvent(h->stream, s.ev_h2d, 0));
  KernelParams kp = make_params(h, s.d_frames, d_nu, stride, s.d_tps, s.d_count, nullptr, nullptr);
  SW_CUDA(h, launch<false>(h, kp, h->stream));
  SW_CUDA(h, cudaEventRecord(s.ev_kernel, h->stream));
  SW_CUDA(h, cudaStreamWaitEvent(s.stream, s.ev_kernel, 0));
  SW_CUDA(h, cudaMemcpyAsync(s.h_count, s.d_count, sizeof(unsigned), cudaMemcpyDeviceToHost, s.stream));
  SW_CUDA(h, cudaEventRecord(s.ev_count, s.stream));
  s.n_ready = s.n_taken = 0;
  s.state.store(kCopying, std::memory_order_release);
  h->counters.units_processed += units;
  h->counters.samples_processed += units * h->channels * h->ticks;
  h->counters.batches++;
  h->next_dispatch++;
  return SWTPG_OK;
}

} // namespace

extern "C" {

uint32_t
swtpg_abi_version(void)
{
  return SWTPG_ABI_VERSION;
}

const char*
swtpg_status_string(swtpg_status s)
{
  switch (s) {
    case SWTPG_OK: return "ok";
    case SWTPG_ERR_INVALID_ARG: return "invalid argument";
    case SWTPG_ERR_CUDA: return "CUDA error / no usable device";
    case SWTPG_ERR_BUSY: return "busy (back-pressure)";
    case SWTPG_ERR_OVERFLOW: return "TP buffer overflow";
    case SWTPG_ERR_STATE: return "invalid call sequence";
    case SWTPG_ERR_UNSUPPORTED: return "unsupported algorithm/format";
  }
  return "unknown";
}

const char*
swtpg_last_error(const swtpg_handle* h)
{
  return h ? h->last_error.c_str() : g_create_error.c_str();
}

int
swtpg_device_available(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return 0;
  }
  for (int d = 0; d < n; ++d) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10)
      return 1;
  }
  return 0;
}

int
swtpg_firwin_int(int n, double cutoff, int multiplier, int16_t* taps)
{ // src/wib2/tpg/DesignFIR.cpp:20-68 (host, double precision; runs once per start in the reference)
  if (n < 2 || n > 64 || !taps)
    return -1;
  const double pi = 3.14159265358979323846;
  std::vector<double> v(size_t(n), 0.0);
  double sum = 0;
  const int alpha = n / 2;
  for (int m = 0; m < n; ++m) {
    const double w = 0.54 - 0.46 * std::cos(2.0 * pi * m / (n - 1));
    const double x = cutoff * (m - alpha);
    v[size_t(m)] = w * (x == 0 ? 1.0 : std::sin(pi * x) / (pi * x));
    sum += v[size_t(m)];
  }
  for (int m = 0; m < n; ++m)
    taps[m] = int16_t(std::round(multiplier * (v[size_t(m)] / sum)));
  return n;
}

swtpg_status
swtpg_create(const swtpg_config* cfg, swtpg_handle** out)
{
  if (!cfg || !out)
    return fail(nullptr, SWTPG_ERR_INVALID_ARG, "cfg/out is NULL");
  *out = nullptr;
  if (cfg->struct_size != sizeof(swtpg_config))
    return fail(nullptr, SWTPG_ERR_INVALID_ARG, "swtpg_config.struct_size mismatch");
  if (cfg->n_links == 0 || cfg->max_units == 0)
    return fail(nullptr, SWTPG_ERR_INVALID_ARG, "n_links and max_units must be > 0");
  if (cfg->max_units >= (1u << kHitUnitBits))
    return fail(nullptr, SWTPG_ERR_INVALID_ARG, "max_units must stay below 2^18 units per batch (hit records carry the unit index in 18 bits)");
  if (cfg->format != SWTPG_FORMAT_WIBETH && cfg->format != SWTPG_FORMAT_WIB2)
    return fail(nullptr, SWTPG_ERR_UNSUPPORTED, "unknown frame format");
  if (cfg->algorithm < SWTPG_ALGO_SIMPLE_THRESHOLD || cfg->algorithm > SWTPG_ALGO_FIR_IQR)
    return fail(nullptr, SWTPG_ERR_UNSUPPORTED, "unknown tpg_algorithm (reference: TPGAlgorithmInexistent)");
  if (cfg->format == SWTPG_FORMAT_WIB2 && cfg->algorithm == SWTPG_ALGO_STANDARD_RS)
    return fail(nullptr, SWTPG_ERR_UNSUPPORTED, "StandardRS does not exist for the WIB2 format (reference: SimpleThreshold and AbsRS only)");
  if (cfg->format == SWTPG_FORMAT_WIB2 && cfg->algorithm == SWTPG_ALGO_ABS_RS && cfg->threshold == 0)
    return fail(nullptr, SWTPG_ERR_INVALID_ARG, "WIB2 AbsRS needs threshold >= 1 (sigmaMax = 2^15 / (multiplier * threshold))");
  if (cfg->wib2_adc_offset % 4 != 0 || cfg->wib2_adc_offset > SWTPG_WIB2_FRAME_BYTES - 448)
    return fail(nullptr, SWTPG_ERR_INVALID_ARG, "wib2_adc_offset must be a multiple of 4 and leave room for the 448-byte ADC block");
  if (cfg->tap_exponent > 14)
    return fail(nullptr, SWTPG_ERR_INVALID_ARG, "tap_exponent out of range");

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    return fail(nullptr, SWTPG_ERR_CUDA, "no CUDA device visible (this library has no CPU fallback)");
  }
  if (cfg->device < 0 || cfg->device >= ndev)
    return fail(nullptr, SWTPG_ERR_INVALID_ARG, "device ordinal out of range");
  SW_CUDA((swtpg_handle*)nullptr, cudaSetDevice(cfg->device));
  int major = 0;
  SW_CUDA((swtpg_handle*)nullptr, cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, cfg->device));
  if (major != 10)
    return fail(nullptr, SWTPG_ERR_CUDA, "device is not compute capability 10.x (kernels are sm_100a only)");

  std::unique_ptr<swtpg_handle> h(new (std::nothrow) swtpg_handle);
  if (!h)
    return fail(nullptr, SWTPG_ERR_INVALID_ARG, "out of host memory");
  h->cfg = *cfg;
  const bool wib2 = cfg->format == SWTPG_FORMAT_WIB2;
  h->unit_bytes = wib2 ? SWTPG_WIB2_SUPERCHUNK_BYTES : SWTPG_WIBETH_FRAME_BYTES;
  h->channels = wib2 ? SWTPG_WIB2_CHANNELS : SWTPG_WIBETH_CHANNELS;
  h->ticks = wib2 ? SWTPG_WIB2_TICKS : SWTPG_WIBETH_TICKS;
  h->groups_per_link = h->channels / 64;
  h->n_groups = cfg->n_links * h->groups_per_link;
  if (h->cfg.n_slots == 0)
    h->cfg.n_slots = 3;
  if (h->cfg.n_slots < 2)
    return fail(nullptr, SWTPG_ERR_INVALID_ARG, "n_slots must be >= 2");
  if (h->cfg.tap_exponent == 0)
    h->cfg.tap_exponent = 6;
  if (h->cfg.wib2_adc_offset == 0)
    h->cfg.wib2_adc_offset = 20;
  bool any_tap = false;
  for (int i = 0; i < 8; ++i)
    any_tap |= h->cfg.fir_taps[i] != 0;
  if (!any_tap) { // src/wib2/WIB2FrameProcessor.cpp:93-94
    swtpg_firwin_int(7, 0.1, 1 << h->cfg.tap_exponent, h->cfg.fir_taps);
    h->cfg.fir_taps[7] = 0;
  }
  // TP buffer: explicit, else samples/128 per batch clamped to [64 Ki, 16 Mi] records (physical TP rates are orders of
  // magnitude below one per 128 samples; the high-occupancy stress configuration sets tp_capacity itself).
  const uint64_t samples = uint64_t(cfg->n_links) * cfg->max_units * h->channels * h->ticks;
  uint64_t cap = cfg->tp_capacity ? cfg->tp_capacity : std::min<uint64_t>(std::max<uint64_t>(samples / 128, 1u << 16), 1u << 24);
  cap = std::min<uint64_t>(cap, std::max<uint64_t>(samples / 2, 1));
  h->tp_capacity = uint32_t(cap);
  // Packed fast path validity (see PackedSimpleWibEth)
  h->fast_simple = cfg->algorithm == SWTPG_ALGO_SIMPLE_THRESHOLD && cfg->threshold <= 32767 &&
                   (wib2 || (cfg->frugal_acc_limit >= 1 && cfg->frugal_acc_limit <= 1000)) && getenv("SWTPG_FORCE_SCALAR") == nullptr;

  h->fast_rs = !wib2 && (cfg->algorithm == SWTPG_ALGO_ABS_RS || cfg->algorithm == SWTPG_ALGO_STANDARD_RS) && cfg->threshold <= 32767 &&
               cfg->frugal_acc_limit >= 1 && cfg->frugal_acc_limit <= 1000 && getenv("SWTPG_FORCE_SCALAR") == nullptr;
  // WIB2 AbsRS fast path validity (see PackedRsIqrWib2); threshold >= 1 is enforced below for the algorithm as such
  if (wib2 && cfg->algorithm == SWTPG_ALGO_ABS_RS && cfg->threshold >= 1) {
    const uint32_t e = h->cfg.tap_exponent, mult = 1u << e;
    const uint64_t sigma_max = (1u << 15) / (uint64_t(mult) * cfg->threshold);
    h->fast_rs_wib2 = e >= 1 && e <= 10 && (sigma_max + 3) * cfg->threshold < 65536 && getenv("SWTPG_FORCE_SCALAR") == nullptr;
  }
  // Packed FIR fast path validity (see PackedFirIqr): binomial taps, and (sigmaMax + 3) * multiplier * threshold < 2^16
  {
    static const int16_t kBinomial[8] = { 1, 6, 15, 20, 15, 6, 1, 0 };
    bool taps_ok = true;
    for (int i = 0; i < 8; ++i)
      taps_ok &= h->cfg.fir_taps[i] == kBinomial[i];
    const uint32_t e = h->cfg.tap_exponent, mult = 1u << e;
    const uint64_t sigma_max = (1u << 15) / (mult * 5u);
    const bool range_ok = cfg->algorithm == SWTPG_ALGO_FIR_IQR && e >= 1 && e <= 10 &&
                          (sigma_max + 3) * mult * uint64_t(cfg->threshold) < 65536 && getenv("SWTPG_FORCE_SCALAR") == nullptr;
    h->fast_fir = range_ok && taps_ok;      // binomial cascade
    h->fast_fir_any = range_ok && !taps_ok; // any other taps[0..6]: packed multiply-add chain (PackedFirIqrAnyTaps)
  }

  swtpg_handle* hp = h.get();
  SW_CUDA(hp, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  SW_CUDA(hp, cudaEventCreate(&h->ev0));
  SW_CUDA(hp, cudaEventCreate(&h->ev1));
  SW_CUDA(hp, cudaMalloc(&h->d_state, size_t(h->n_groups) * kStateWordsPerGroup * 4));
  SW_CUDA(hp, cudaMalloc(&h->d_flags, size_t(h->n_groups) * 4));
  SW_CUDA(hp, cudaMalloc(&h->d_link_cursor, 2 * sizeof(uint32_t)));
  SW_CUDA(hp, cudaMemset(h->d_link_cursor, 0, 2 * sizeof(uint32_t)));
  SW_CUDA(hp, cudaMalloc(&h->d_tps, size_t(h->tp_capacity) * sizeof(swtpg_tp)));
  SW_CUDA(hp, cudaMalloc(&h->d_count, sizeof(unsigned)));
  SW_CUDA(hp, cudaMallocHost(&h->h_count, sizeof(unsigned)));
  SW_CUDA(hp, cudaMalloc(&h->d_nunits, size_t(cfg->n_links) * 4));
  SW_CUDA(hp, cudaMallocHost(&h->h_nunits, size_t(cfg->n_links) * 4));
  h->submitted.reset(new std::atomic<uint64_t>[cfg->n_links]);
  h->link_range.reset(new swtpg_handle::HostRange[cfg->n_links]);
  h->link_range_epoch.reset(new uint64_t[cfg->n_links]());
  for (uint32_t l = 0; l < cfg->n_links; ++l)
    h->submitted[l].store(0);
  *out = h.release();
  return SWTPG_OK;
}

void
swtpg_destroy(swtpg_handle* h)
{
  if (!h)
    return;
  cudaSetDevice(h->cfg.device);
  cudaDeviceSynchronize();
  for (const auto& r : h->ranges)
    if (cudaHostUnregister(reinterpret_cast<void*>(r.lo)) != cudaSuccess)
      cudaGetLastError();
  for (auto& s : h->slots)
    free_slot(*s);
  for (auto& b : h->bounce) {
    for (int i = 0; i < 2; ++i) {
      if (b.buf[i]) cudaFreeHost(b.buf[i]);
      if (b.free_ev[i]) cudaEventDestroy(b.free_ev[i]);
    }
    if (b.done) cudaEventDestroy(b.done);
    if (b.stream) cudaStreamDestroy(b.stream);
  }
  if (h->d_state) cudaFree(h->d_state);
  if (h->d_flags) cudaFree(h->d_flags);
  if (h->d_link_cursor) cudaFree(h->d_link_cursor);
  if (h->d_tps) cudaFree(h->d_tps);
  if (h->d_count) cudaFree(h->d_count);
  if (h->h_count) cudaFreeHost(h->h_count);
  if (h->d_nunits) cudaFree(h->d_nunits);
  if (h->h_nunits) cudaFreeHost(h->h_nunits);
  if (h->d_frames) cudaFree(h->d_frames);
  if (h->d_ped) cudaFree(h->d_ped);
  if (h->d_wav) cudaFree(h->d_wav);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete[] h->h_rs_factor;
  delete h;
}

swtpg_status
swtpg_start(swtpg_handle* h)
{
  if (!h)
    return SWTPG_ERR_INVALID_ARG;
  SW_CUDA(h, cudaSetDevice(h->cfg.device));
  swtpg_status s = reset_state(h);
  if (s != SWTPG_OK)
    return s;
  for (uint32_t l = 0; l < h->cfg.n_links; ++l)
    h->submitted[l].store(0);
  for (size_t i = 0; i < h->slots.size(); ++i) {
    Slot& sl = *h->slots[i];
    sl.batch.store(i);
    sl.remaining.store(h->cfg.n_links * h->cfg.max_units);
    sl.n_borrowed.store(0);
    sl.state.store(kFilling);
  }
  h->next_dispatch = h->next_poll = 0;
  h->counters = swtpg_counters{};
  h->submit_busy.store(0);
  h->timed = false;
  h->started = true;
  return SWTPG_OK;
}

swtpg_status
swtpg_stop(swtpg_handle* h)
{
  if (!h)
    return SWTPG_ERR_INVALID_ARG;
  swtpg_status s = swtpg_sync(h);
  h->started = false;
  return s;
}

swtpg_status
swtpg_set_rs_memory_factor(swtpg_handle* h, const uint16_t* by_link_channel)
{
  if (!h)
    return SWTPG_ERR_INVALID_ARG;
  delete[] h->h_rs_factor;
  h->h_rs_factor = nullptr;
  if (by_link_channel) {
    const size_t n = size_t(h->cfg.n_links) * h->channels;
    h->h_rs_factor = new uint16_t[n];
    memcpy(h->h_rs_factor, by_link_channel, n * sizeof(uint16_t));
  }
  if (h->started) { // only the factor row changes; carried state is preserved
    SW_CUDA(h, cudaSetDevice(h->cfg.device));
    std::vector<uint32_t> row(32);
    for (uint32_t g = 0; g < h->n_groups; ++g) {
      const uint32_t link = g / h->groups_per_link, sub = g % h->groups_per_link;
      for (uint32_t lane = 0; lane < 32; ++lane) {
        uint32_t lo = h->cfg.rs_memory_factor, hi = lo;
        if (h->h_rs_factor) {
          const uint16_t* f = h->h_rs_factor + size_t(link) * h->channels + sub * 64 + 2 * lane;
          lo = f[0];
          hi = f[1];
        }
        row[lane] = lo | (hi << 16);
      }
      SW_CUDA(h, cudaMemcpy(h->d_state + size_t(g) * kStateWordsPerGroup + SV_RS_FACTOR * 32, row.data(), 128, cudaMemcpyHostToDevice));
    }
  }
  return SWTPG_OK;
}

swtpg_status
swtpg_process_device(swtpg_handle* h, const void* d_frames, const uint32_t* n_units, uint32_t units_stride, void* stream)
{
  swtpg_status s = check_batch_args(h, d_frames, n_units, units_stride);
  if (s != SWTPG_OK)
    return s;
  if (reinterpret_cast<uintptr_t>(d_frames) & 15u)
    return fail(h, SWTPG_ERR_INVALID_ARG, "d_frames must be 16-byte aligned (bulk-copy source)");
  SW_CUDA(h, cudaSetDevice(h->cfg.device));
  cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : h->stream;
  return enqueue_batch(h, d_frames, n_units, units_stride, st, false);
}

swtpg_status
swtpg_fetch_tps(swtpg_handle* h, swtpg_tp* out, size_t cap, size_t* n_out)
{
  if (!h)
    return SWTPG_ERR_INVALID_ARG;
  if (!h->last_stream)
    return fail(h, SWTPG_ERR_STATE, "no batch has been processed");
  SW_CUDA(h, cudaSetDevice(h->cfg.device));
  return fetch(h, h->last_stream, out, cap, n_out);
}

double
swtpg_last_kernel_ms(swtpg_handle* h)
{
  if (!h || !h->timed)
    return -1.0;
  if (cudaEventSynchronize(h->ev1) != cudaSuccess)
    return -1.0;
  float ms = 0;
  if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) != cudaSuccess)
    return -1.0;
  return double(ms);
}

// Copy of one payload into pinned memory (csrc/stage_copy.cpp: non-temporal stores where the CPU has AVX2).
extern "C" void swtpg_stage_copy(void* dst, const void* src, size_t bytes);

// Host-to-device copy of a PAGEABLE source. cudaMemcpyAsync would stage it through the driver's own bounce buffer on one
// thread (11 GB/s on the bench box against 55 GB/s from pinned memory); here a few worker threads copy 4 MB chunks into their
// own pinned double buffers with non-temporal stores and queue the DMA of each chunk on their own stream, so the host copies
// of later chunks overlap the transfers of earlier ones. The handle's stream then waits for every worker's last transfer.
constexpr size_t kBounceChunk = size_t(4) << 20;
constexpr size_t kBounceMinBytes = size_t(16) << 20;

static bool
is_pageable(const void* p)
{
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}

static swtpg_status
staged_h2d(swtpg_handle* h, uint8_t* d_dst, const uint8_t* src, size_t bytes)
{
  if (h->bounce.empty()) {
    const unsigned hw = std::max(2u, std::thread::hardware_concurrency());
    const unsigned workers = std::min(8u, hw / 2);
    std::vector<swtpg_handle::Bounce> b(workers);
    for (auto& w : b) {
      for (int i = 0; i < 2; ++i) {
        SW_CUDA(h, cudaMallocHost(&w.buf[i], kBounceChunk));
        SW_CUDA(h, cudaEventCreateWithFlags(&w.free_ev[i], cudaEventDisableTiming));
      }
      SW_CUDA(h, cudaEventCreateWithFlags(&w.done, cudaEventDisableTiming));
      SW_CUDA(h, cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking));
    }
    h->bounce = std::move(b);
  }
  const size_t n_chunks = (bytes + kBounceChunk - 1) / kBounceChunk;
  const size_t workers = h->bounce.size();
  std::vector<cudaError_t> err(workers, cudaSuccess);
  std::vector<std::thread> threads;
  const int device = h->cfg.device;
  for (size_t t = 0; t < workers; ++t)
    threads.emplace_back([&, t]() {
      swtpg_handle::Bounce& w = h->bounce[t];
      cudaError_t e = cudaSetDevice(device);
      unsigned used = 0;
      for (size_t c = t; c < n_chunks && e == cudaSuccess; c += workers, ++used) {
        const int slot = int(used & 1u);
        if (used >= 2)
          e = cudaEventSynchronize(w.free_ev[slot]); // the transfer that last read this buffer has finished
        if (e != cudaSuccess)
          break;
        const size_t off = c * kBounceChunk, len = std::min(kBounceChunk, bytes - off);
        swtpg_stage_copy(w.buf[slot], src + off, len);
        e = cudaMemcpyAsync(d_dst + off, w.buf[slot], len, cudaMemcpyHostToDevice, w.stream);
        if (e == cudaSuccess)
          e = cudaEventRecord(w.free_ev[slot], w.stream);
      }
      if (e == cudaSuccess)
        e = cudaEventRecord(w.done, w.stream);
      err[t] = e;
    });
  for (auto& th : threads)
    th.join();
  for (size_t t = 0; t < workers; ++t) {
    SW_CUDA(h, err[t]);
    SW_CUDA(h, cudaStreamWaitEvent(h->stream, h->bounce[t].done, 0));
  }
  return SWTPG_OK;
}

static swtpg_status
process_host_impl(swtpg_handle* h, const void* frames, const uint32_t* n_units, uint32_t stride, swtpg_tp* out, size_t cap, size_t* n_out,
                  int16_t* ped_out, int16_t* wav_out, bool dump)
{
  swtpg_status s = check_batch_args(h, frames, n_units, stride);
  if (s != SWTPG_OK)
    return s;
  SW_CUDA(h, cudaSetDevice(h->cfg.device));
  const size_t bytes = size_t(h->cfg.n_links) * stride * h->unit_bytes;
  if (bytes > h->d_frames_bytes) {
    if (h->d_frames)
      cudaFree(h->d_frames);
    h->d_frames = nullptr;
    h->d_frames_bytes = 0;
    SW_CUDA(h, cudaMalloc(&h->d_frames, bytes));
    h->d_frames_bytes = bytes;
  }
  const size_t dump_elems = size_t(h->cfg.n_links) * stride * h->ticks * h->channels;
  if (dump && dump_elems > h->d_dump_elems) {
    if (h->d_ped) cudaFree(h->d_ped);
    if (h->d_wav) cudaFree(h->d_wav);
    h->d_ped = h->d_wav = nullptr;
    h->d_dump_elems = 0;
    SW_CUDA(h, cudaMalloc(&h->d_ped, dump_elems * 2));
    SW_CUDA(h, cudaMalloc(&h->d_wav, dump_elems * 2));
    h->d_dump_elems = dump_elems;
  }
  if (dump && dump_elems) {
    SW_CUDA(h, cudaMemsetAsync(h->d_ped, 0, dump_elems * 2, h->stream));
    SW_CUDA(h, cudaMemsetAsync(h->d_wav, 0, dump_elems * 2, h->stream));
  }
  if (bytes >= kBounceMinBytes && is_pageable(frames)) {
    s = staged_h2d(h, h->d_frames, static_cast<const uint8_t*>(frames), bytes);
    if (s != SWTPG_OK)
      return s;
  } else if (bytes) {
    SW_CUDA(h, cudaMemcpyAsync(h->d_frames, frames, bytes, cudaMemcpyHostToDevice, h->stream));
  }
  h->counters.h2d_bytes += bytes;
  s = enqueue_batch(h, h->d_frames, n_units, stride, h->stream, dump);
  if (s != SWTPG_OK)
    return s;
  s = fetch(h, h->stream, out, cap, n_out);
  if (dump && dump_elems) {
    if (ped_out)
      SW_CUDA(h, cudaMemcpy(ped_out, h->d_ped, dump_elems * 2, cudaMemcpyDeviceToHost));
    if (wav_out)
      SW_CUDA(h, cudaMemcpy(wav_out, h->d_wav, dump_elems * 2, cudaMemcpyDeviceToHost));
  }
  return s;
}

swtpg_status
swtpg_process_host(swtpg_handle* h, const void* frames, const uint32_t* n_units, uint32_t units_stride, swtpg_tp* out, size_t cap,
                   size_t* n_out)
{
  return process_host_impl(h, frames, n_units, units_stride, out, cap, n_out, nullptr, nullptr, false);
}

swtpg_status
swtpg_process_host_debug(swtpg_handle* h, const void* frames, const uint32_t* n_units, uint32_t units_stride, swtpg_tp* out, size_t cap,
                         size_t* n_out, int16_t* pedestal_out, int16_t* waveform_out)
{
  return process_host_impl(h, frames, n_units, units_stride, out, cap, n_out, pedestal_out, waveform_out, true);
}

// ---- streaming path ---------------------------------------------------------------------------------------------------
// Is [unit, unit + bytes) inside a range registered with swtpg_register_buffer? Lock-free on the hot path: every link's
// producer thread keeps the last range it hit (a link's payloads come from one latency buffer).
static bool
is_registered(swtpg_handle* h, uint32_t link, const void* unit, size_t bytes)
{
  const uint64_t epoch = h->ranges_epoch.load(std::memory_order_acquire);
  if (epoch == 0)
    return false; // nothing was ever registered
  const uintptr_t a = reinterpret_cast<uintptr_t>(unit);
  swtpg_handle::HostRange& c = h->link_range[link];
  if (h->link_range_epoch[link] == epoch) {
    if (a >= c.lo && a + bytes <= c.hi)
      return true;
    if (c.hi == 0)
      return false; // cached miss: this link's payloads are not in registered memory
  }
  std::lock_guard<std::mutex> lk(h->ranges_mu); // first payload of the link, or the set of ranges changed, or another range
  c = swtpg_handle::HostRange{};
  for (const auto& r : h->ranges)
    if (a >= r.lo && a + bytes <= r.hi)
      c = r;
  h->link_range_epoch[link] = h->ranges_epoch.load(std::memory_order_relaxed);
  return c.hi != 0;
}

swtpg_status
swtpg_submit(swtpg_handle* h, uint32_t link, const void* unit, size_t bytes)
{
  if (!h || !unit)
    return SWTPG_ERR_INVALID_ARG;
  if (!h->started)
    return fail(h, SWTPG_ERR_STATE, "swtpg_start has not been called");
  if (link >= h->cfg.n_links || bytes != h->unit_bytes)
    return fail(h, SWTPG_ERR_INVALID_ARG, "bad link index or unit size");
  if (!h->slots_ready.load(std::memory_order_acquire)) {
    std::lock_guard<std::mutex> lk(h->dispatch_mu);
    SW_CUDA(h, cudaSetDevice(h->cfg.device));
    swtpg_status s = ensure_slots(h);
    if (s != SWTPG_OK)
      return s;
  }
  const uint64_t seq = h->submitted[link].load(std::memory_order_relaxed); // one producer thread per link
  const uint64_t batch = seq / h->cfg.max_units;
  const uint32_t u = uint32_t(seq % h->cfg.max_units);
  Slot& s = *h->slots[batch % h->slots.size()];
  if (s.batch.load(std::memory_order_acquire) != batch || s.state.load(std::memory_order_acquire) != kFilling) {
    h->submit_busy.fetch_add(1, std::memory_order_relaxed);
    return SWTPG_ERR_BUSY; // ring full: the caller drops or retries, like a failed try_send
  }
  const size_t idx = size_t(link) * h->cfg.max_units + u;
  if (is_registered(h, link, unit, bytes)) { // zero-copy: the batch's H2D reads the unit where it lies
    s.borrowed[idx] = static_cast<const uint8_t*>(unit);
    s.n_borrowed.fetch_add(1, std::memory_order_relaxed);
  } else {
    s.borrowed[idx] = nullptr;
    swtpg_stage_copy(s.h_frames + idx * h->unit_bytes, unit, bytes);
  }
  h->submitted[link].store(seq + 1, std::memory_order_release);
  if (s.remaining.fetch_sub(1, std::memory_order_acq_rel) == 1) { // this unit completed the batch
    std::lock_guard<std::mutex> lk(h->dispatch_mu);
    SW_CUDA(h, cudaSetDevice(h->cfg.device));
    return dispatch_slot(h, s, nullptr);
  }
  return SWTPG_OK;
}

swtpg_status
swtpg_register_buffer(swtpg_handle* h, void* base, size_t bytes)
{
  if (!h || !base || bytes == 0)
    return SWTPG_ERR_INVALID_ARG;
  SW_CUDA(h, cudaSetDevice(h->cfg.device));
  cudaError_t e = cudaHostRegister(base, bytes, cudaHostRegisterPortable);
  if (e == cudaErrorHostMemoryAlreadyRegistered)
    cudaGetLastError(); // e.g. two handles (GPUs) sharing one latency buffer: fine, it is pinned
  else
    SW_CUDA(h, e);
  std::lock_guard<std::mutex> lk(h->ranges_mu);
  const uintptr_t lo = reinterpret_cast<uintptr_t>(base);
  h->ranges.push_back({ lo, lo + bytes });
  h->ranges_epoch.fetch_add(1, std::memory_order_release);
  return SWTPG_OK;
}

swtpg_status
swtpg_unregister_buffer(swtpg_handle* h, void* base)
{
  if (!h || !base)
    return SWTPG_ERR_INVALID_ARG;
  swtpg_status st = swtpg_sync(h); // no copy engine may still be reading from it
  if (st != SWTPG_OK)
    return st;
  {
    std::lock_guard<std::mutex> lk(h->ranges_mu);
    const uintptr_t lo = reinterpret_cast<uintptr_t>(base);
    auto it = std::find_if(h->ranges.begin(), h->ranges.end(), [lo](const swtpg_handle::HostRange& r) { return r.lo == lo; });
    if (it == h->ranges.end())
      return fail(h, SWTPG_ERR_INVALID_ARG, "buffer was not registered with this handle");
    h->ranges.erase(it);
    h->ranges_epoch.fetch_add(1, std::memory_order_release);
  }
  cudaError_t e = cudaHostUnregister(base);
  if (e != cudaSuccess)
    cudaGetLastError(); // registered by another handle first, or already gone
  return SWTPG_OK;
}

swtpg_status
swtpg_flush(swtpg_handle* h)
{
  if (!h)
    return SWTPG_ERR_INVALID_ARG;
  if (!h->started)
    return fail(h, SWTPG_ERR_STATE, "swtpg_start has not been called");
  if (!h->slots_ready.load(std::memory_order_acquire))
    return SWTPG_OK;
  std::lock_guard<std::mutex> lk(h->dispatch_mu);
  SW_CUDA(h, cudaSetDevice(h->cfg.device));
  // Must not race with swtpg_submit. Closes the oldest partially filled batch (and any later one that links running
  // ahead already started), then realigns every link to the next batch boundary.
  for (;;) {
    const uint64_t b = h->next_dispatch;
    Slot& s = *h->slots[b % h->slots.size()];
    if (s.batch.load() != b || s.state.load() != kFilling)
      break;
    std::vector<uint32_t> nu(h->cfg.n_links);
    uint64_t total = 0;
    for (uint32_t l = 0; l < h->cfg.n_links; ++l) {
      const uint64_t seq = h->submitted[l].load();
      const uint64_t lo = b * h->cfg.max_units;
      nu[l] = seq <= lo ? 0u : uint32_t(std::min<uint64_t>(seq - lo, h->cfg.max_units));
      total += nu[l];
    }
    if (total == 0)
      break;
    for (uint32_t l = 0; l < h->cfg.n_links; ++l)
      if (h->submitted[l].load() < (b + 1) * h->cfg.max_units)
        h->submitted[l].store((b + 1) * h->cfg.max_units);
    swtpg_status st = dispatch_slot(h, s, nu.data());
    if (st != SWTPG_OK)
      return st;
  }
  return SWTPG_OK;
}

swtpg_status
swtpg_poll(swtpg_handle* h, swtpg_tp* out, size_t cap, size_t* n_out)
{
  if (n_out)
    *n_out = 0;
  if (!h || (!out && cap))
    return SWTPG_ERR_INVALID_ARG;
  if (!h->slots_ready.load(std::memory_order_acquire))
    return SWTPG_OK;
  std::lock_guard<std::mutex> lk(h->dispatch_mu);
  SW_CUDA(h, cudaSetDevice(h->cfg.device));
  size_t n = 0;
  swtpg_status ret = SWTPG_OK;
  for (;;) {
    Slot& s = *h->slots[h->next_poll % h->slots.size()];
    if (s.batch.load() != h->next_poll)
      break;
    int st = s.state.load(std::memory_order_acquire);
    if (st == kCopying) {
      if (cudaEventQuery(s.ev_count) != cudaSuccess) {
        cudaGetLastError();
        break;
      }
      const unsigned found = *s.h_count;
      const unsigned stored = std::min<unsigned>(found, h->tp_capacity);
      h->counters.tps_emitted += found;
      if (found > stored) {
        h->counters.tps_dropped_overflow += found - stored;
        ret = SWTPG_ERR_OVERFLOW;
        h->last_error = "device TP buffer overflow: raise swtpg_config.tp_capacity";
      }
      s.n_ready = stored;
      s.n_taken = 0;
      if (stored)
        SW_CUDA(h, cudaMemcpyAsync(s.h_tps, s.d_tps, size_t(stored) * sizeof(swtpg_tp), cudaMemcpyDeviceToHost, s.stream));
      SW_CUDA(h, cudaEventRecord(s.ev_tps, s.stream));
      h->counters.d2h_bytes += size_t(stored) * sizeof(swtpg_tp) + sizeof(unsigned);
      s.state.store(kFetching);
      st = kFetching;
    }
    if (st == kFetching) {
      if (cudaEventQuery(s.ev_tps) != cudaSuccess) {
        cudaGetLastError();
        break;
      }
      s.state.store(kReady);
      st = kReady;
    }
    if (st != kReady)
      break;
    const uint32_t take = uint32_t(std::min<size_t>(cap - n, s.n_ready - s.n_taken));
    if (take)
      memcpy(out + n, s.h_tps + s.n_taken, size_t(take) * sizeof(swtpg_tp));
    n += take;
    s.n_taken += take;
    if (s.n_taken < s.n_ready)
      break; // caller's buffer is full; the rest comes with the next poll
    // recycle the slot for batch next_poll + n_slots
    s.remaining.store(h->cfg.n_links * h->cfg.max_units);
    s.n_borrowed.store(0, std::memory_order_relaxed);
    s.batch.store(h->next_poll + h->slots.size(), std::memory_order_release);
    s.state.store(kFilling, std::memory_order_release);
    h->next_poll++;
  }
  if (n_out)
    *n_out = n;
  return ret;
}

swtpg_status
swtpg_sync(swtpg_handle* h)
{
  if (!h)
    return SWTPG_ERR_INVALID_ARG;
  SW_CUDA(h, cudaSetDevice(h->cfg.device));
  SW_CUDA(h, cudaStreamSynchronize(h->stream));
  for (auto& s : h->slots)
    SW_CUDA(h, cudaStreamSynchronize(s->stream));
  if (h->last_stream && h->last_stream != h->stream)
    SW_CUDA(h, cudaStreamSynchronize(h->last_stream));
  return SWTPG_OK;
}

swtpg_status
swtpg_dump_state(swtpg_handle* h, uint32_t link, swtpg_channel_state* out)
{
  if (!h || !out || link >= h->cfg.n_links)
    return SWTPG_ERR_INVALID_ARG;
  SW_CUDA(h, cudaSetDevice(h->cfg.device));
  swtpg_status s = swtpg_sync(h);
  if (s != SWTPG_OK)
    return s;
  std::vector<uint32_t> st(size_t(h->groups_per_link) * kStateWordsPerGroup);
  std::vector<uint32_t> fl(h->groups_per_link);
  SW_CUDA(h, cudaMemcpy(st.data(), h->d_state + size_t(link) * h->groups_per_link * kStateWordsPerGroup, st.size() * 4, cudaMemcpyDeviceToHost));
  SW_CUDA(h, cudaMemcpy(fl.data(), h->d_flags + size_t(link) * h->groups_per_link, fl.size() * 4, cudaMemcpyDeviceToHost));
  for (uint32_t c = 0; c < h->channels; ++c) {
    const uint32_t g = c / 64, lane = (c % 64) / 2, hi = c & 1;
    auto get = [&](int v) -> uint16_t {
      const uint32_t w = st[size_t(g) * kStateWordsPerGroup + size_t(v) * 32 + lane];
      return uint16_t(hi ? (w >> 16) : (w & 0xFFFFu));
    };
    swtpg_channel_state& o = out[c];
    memset(&o, 0, sizeof o);
    o.pedestal = int16_t(get(SV_MEDIAN));
    o.accum = int16_t(get(SV_ACCUM));
    o.quantile25 = int16_t(get(SV_Q25));
    o.quantile75 = int16_t(get(SV_Q75));
    o.accum25 = int16_t(get(SV_A25));
    o.accum75 = int16_t(get(SV_A75));
    o.rs = int16_t(get(SV_RS));
    o.pedestal_rs = int16_t(get(SV_MED_RS));
    o.accum_rs = int16_t(get(SV_ACC_RS));
    o.rs_memory_factor = get(SV_RS_FACTOR);
    o.prev_was_over = get(SV_PREV);
    o.hit_charge = get(SV_CHARGE);
    o.hit_tover = get(SV_TOVER);
    o.hit_peak_adc = get(SV_PEAK_ADC);
    o.hit_peak_time = get(SV_PEAK_TIME);
    o.initialized = uint16_t(fl[g] & kFlagInitialized);
    for (int j = 0; j < 8; ++j)
      o.prev_samp[j] = int16_t(get(SV_RING0 + j));
  }
  return SWTPG_OK;
}

swtpg_status
swtpg_get_counters(swtpg_handle* h, swtpg_counters* out)
{
  if (!h || !out)
    return SWTPG_ERR_INVALID_ARG;
  *out = h->counters;
  out->submit_busy = h->submit_busy.load();
  return SWTPG_OK;
}

void*
swtpg_alloc_pinned(size_t bytes, int write_combined)
{
  void* p = nullptr;
  const unsigned flags = cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0u);
  if (bytes == 0 || cudaHostAlloc(&p, bytes, flags) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

void
swtpg_free_pinned(void* p)
{
  if (p && cudaFreeHost(p) != cudaSuccess)
    cudaGetLastError();
}

static inline bool
tp_less(const swtpg_tp& a, const swtpg_tp& b)
{
  if (a.time_start != b.time_start) return a.time_start < b.time_start;
  if (a.link != b.link) return a.link < b.link;
  if (a.channel != b.channel) return a.channel < b.channel;
  if (a.time_over_threshold != b.time_over_threshold) return a.time_over_threshold < b.time_over_threshold;
  return a.adc_integral < b.adc_integral;
}

// Host-side ordering of a batch's TP list. A 64-frame batch of 148 APAs carries ~0.5 M records: std::sort needs ~110 ms for
// them, more than the batch's whole host-to-device copy. The keys are narrow, though — time_start spans the batch (a few
// 10^5 ticks), link < n_links, channel < 256 — so (time_start - min, link, channel) packs into one 64-bit integer, and an LSD
// radix sort of (key, index) pairs with 11-bit digits followed by one gather orders the list in ~10 ms. Ties on the key
// (which cannot come out of one handle) are ordered like tp_less afterwards; lists whose keys do not fit fall back to std::sort.
static unsigned
bit_length(uint64_t v)
{
  unsigned b = 0;
  while (v) {
    ++b;
    v >>= 1;
  }
  return b;
}

static void
sort_tps_impl(swtpg_tp* a, size_t n)
{
  if (n < 2)
    return;
  if (n < 4096 || n > 0xFFFFFFFFull) {
    std::stable_sort(a, a + n, tp_less);
    return;
  }
  uint64_t tmin = a[0].time_start, tmax = a[0].time_start;
  uint32_t lmax = 0;
  uint16_t cmax = 0;
  for (size_t i = 0; i < n; ++i) {
    tmin = std::min(tmin, a[i].time_start);
    tmax = std::max(tmax, a[i].time_start);
    lmax = std::max(lmax, a[i].link);
    cmax = std::max(cmax, a[i].channel);
  }
  const unsigned cb = bit_length(cmax), lb = bit_length(lmax), tb = bit_length(tmax - tmin), bits = cb + lb + tb;
  if (bits > 64) {
    std::stable_sort(a, a + n, tp_less);
    return;
  }
  struct KV
  {
    uint64_t key;
    uint32_t idx, pad;
  };
  constexpr unsigned kDigit = 11, kBuckets = 1u << kDigit, kMaxPasses = (64 + kDigit - 1) / kDigit;
  const unsigned passes = (bits + kDigit - 1) / kDigit;
  // scratch is kept per calling thread between calls (grow-only): a batch-sized sort per superchunk would otherwise spend most
  // of its time faulting in 64 n bytes of fresh pages
  thread_local std::vector<unsigned char> scratch;
  const size_t need = 2 * n * sizeof(KV) + n * sizeof(swtpg_tp) + 64;
  if (scratch.size() < need)
    scratch.resize(need + need / 4);
  unsigned char* base = scratch.data() + ((64 - (reinterpret_cast<uintptr_t>(scratch.data()) & 63)) & 63);
  KV* kv = reinterpret_cast<KV*>(base);
  KV* kv2 = kv + n;
  swtpg_tp* out = reinterpret_cast<swtpg_tp*>(kv2 + n);
  std::vector<uint32_t> hist(size_t(kMaxPasses) * kBuckets, 0u); // all digit histograms in the pass that builds the keys
  for (size_t i = 0; i < n; ++i) {
    const uint64_t t = a[i].time_start - tmin;
    const uint64_t key = (tb ? t << (lb + cb) : 0) | (uint64_t(a[i].link) << cb) | a[i].channel;
    kv[i].key = key;
    kv[i].idx = uint32_t(i);
    for (unsigned p = 0; p < passes; ++p)
      ++hist[p * kBuckets + ((key >> (p * kDigit)) & (kBuckets - 1))];
  }
  KV *src = kv, *dst = kv2;
  for (unsigned p = 0; p < passes; ++p) {
    uint32_t* h = hist.data() + size_t(p) * kBuckets;
    uint32_t sum = 0;
    bool trivial = false;
    for (unsigned d = 0; d < kBuckets; ++d) {
      trivial |= h[d] == n; // every key has the same digit: nothing to do
      const uint32_t c = h[d];
      h[d] = sum;
      sum += c;
    }
    if (trivial)
      continue;
    const unsigned shift = p * kDigit;
    for (size_t i = 0; i < n; ++i)
      dst[h[(src[i].key >> shift) & (kBuckets - 1)]++] = src[i];
    std::swap(src, dst);
  }
  for (size_t i = 0; i < n; ++i)
    out[i] = a[src[i].idx];
  for (size_t i = 0; i < n;) { // runs of equal (time_start, link, channel): order the rest of tp_less, keeping input order on full ties
    size_t j = i + 1;
    while (j < n && src[j].key == src[i].key)
      ++j;
    if (j - i > 1)
      std::stable_sort(out + i, out + j, tp_less);
    i = j;
  }
  memcpy(a, out, n * sizeof(swtpg_tp));
}

void
swtpg_sort_tps(swtpg_tp* tps, size_t n)
{
  if (tps && n > 1)
    sort_tps_impl(tps, n);
}

void
swtpg_merge_sorted(const swtpg_tp* const* lists, const size_t* n, size_t k, swtpg_tp* out)
{
  // The lists are sorted like swtpg_sort_tps; ties are resolved by list index (stable across GPUs). Concatenating them in list
  // order and running the stable radix sort gives exactly that order, in O(total) instead of O(total log k) comparisons.
  size_t total = 0;
  for (size_t i = 0; i < k; ++i) {
    if (n[i])
      memcpy(out + total, lists[i], n[i] * sizeof(swtpg_tp));
    total += n[i];
  }
  sort_tps_impl(out, total);
}

} // extern "C"
