// Host runtime behind include/swtpg.h: device buffers, kernel selection and the batch entry points. The streaming path
// (swtpg_submit ... swtpg_poll) lives in swtpg_stream.cu. Plain CUDA runtime; no torch, no CPU compute fallback.
#include "swtpg_handle.h"
#include "swtpg_kernels.cuh"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

using namespace swtpg;
using swtpg_internal::fail;

// Layouts the ctypes binding (fdreadoutlibs_b200/_lib.py, frames.py) relies on.
static_assert(sizeof(swtpg_tp) == 32, "swtpg_tp must stay 32 bytes (two 16-byte device stores)");
static_assert(sizeof(swtpg_config) == 72, "swtpg_config layout changed: bump SWTPG_ABI_VERSION and the bindings");
static_assert(sizeof(swtpg_channel_state) == 48, "swtpg_channel_state layout changed");
static_assert(sizeof(swtpg_counters) == 80, "swtpg_counters layout changed");

namespace {
thread_local std::string g_create_error;
thread_local std::string g_error_copy;
}
void
swtpg_internal::set_create_error(const char* msg)
{
  g_create_error = msg;
}

namespace {

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
#ifndef SWTPG_GEO_STAGES
#define SWTPG_GEO_STAGES 2
#define SWTPG_GEO_CHUNK 32
#endif
#ifndef SWTPG_GEO_MINCTAS
#define SWTPG_GEO_MINCTAS 1
#endif
using GeoDefault = Geo<1, SWTPG_GEO_STAGES, SWTPG_GEO_CHUNK, SWTPG_GEO_MINCTAS>;

#ifndef SWTPG_AUTO_SLICE_GEOM
#define SWTPG_AUTO_SLICE_GEOM 1 // sliced launches: equal slices (0) or halving ones (1)
#endif

// WARPS_PER_SM: 0 = the policy's own measured optimum (Algo::kWarpsPerSm). SLICE_WHOLE_ROUNDS: hand the links out in slices
// even when their number is a whole multiple of the persistent warps.
template<class Algo, bool DUMP, class G, int WARPS_PER_SM = 0, bool SLICE_WHOLE_ROUNDS = false>
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
  constexpr int kWarpsPerSm = WARPS_PER_SM > 0 ? WARPS_PER_SM : Algo::kWarpsPerSm;
  if (G::warps == 1 && kWarpsPerSm > 0) // the measured optimum, if the device holds that many
    even = std::min<unsigned>(even, unsigned(kWarpsPerSm));
  unsigned warps = std::min<unsigned>(kp.n_links, even * unsigned(sm_count[dev]) * G::warps);
  static const int warps_override = [] { const char* e = getenv("SWTPG_WARPS"); return e ? atoi(e) : 0; }(); // tuning aid
  if (warps_override > 0)
    warps = std::min<unsigned>(unsigned(warps_override), kp.n_links);
  const unsigned grid = std::min<unsigned>((warps + G::warps - 1) / G::warps, unsigned(resident[dev]));
  // More links than persistent warps = several rounds of links per warp, and whatever the last round leaves idle is lost (round 2,
  // straight-line SimpleThreshold policy: 6000 links on 4144 warps 64.9 % of the HBM peak, 8288 = two full rounds 73.0 %). Handing
  // the links out in 4 slices makes the rounds short, at the price of the extra state round trips (about 3 % per launch):
  //   * HALVING slices (n/2, n/4, n/8, n/8 units): what the launch's last round leaves idle is at most one LAST slice, so small
  //     last slices shorten the tail without more round trips per link than four equal ones (pipelined SimpleThreshold policy at
  //     20 warps per SM, 5920 links: whole links 68.9 %, equal slices 70.8 %, halving 73.1 %; 6000 links 71.0 -> 73.3 %; AbsRS + 2.5 %;
  //     profiles/r02_halving_slices_probe.txt). The straight-line policy at 28 warps per SM is the one case that loses with them.
  //   * For most policies a link count that is a whole number of rounds keeps whole links (straight-line form, 8288 links: 73.0 %
  //     against 70.8 % sliced); the pipelined SimpleThreshold policy gains from slices even then (SLICE_WHOLE_ROUNDS).
  // A launch that cannot fill the GPU keeps whole links: time is sequential per link, slices could only wait for each other.
  KernelParams kq = kp;
  unsigned parts = 1;
  const unsigned persistent = grid * unsigned(G::warps);
  if (kp.n_links > persistent && (SLICE_WHOLE_ROUNDS || kp.n_links % persistent != 0))
    parts = std::max(1u, std::min(4u, kp.units_stride / 16u));
  static const int parts_override = [] { const char* e = getenv("SWTPG_PARTS"); return e ? atoi(e) : 0; }(); // tuning aid
  if (parts_override > 0)
    parts = std::min(unsigned(parts_override), 8u);
  kq.parts_log2 = 0;
  while ((2u << kq.parts_log2) <= parts) // the largest power of two not above it
    ++kq.parts_log2;
  static const int geom = [] { const char* e = getenv("SWTPG_SLICE_GEOM"); return e ? atoi(e) : SWTPG_AUTO_SLICE_GEOM; }(); // tuning aid
  kq.slice_geom = geom != 0 ? 1u : 0u;
  k<<<grid, G::warps * 32, G::smem, s>>>(kq);
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
// WIBEth SimpleThreshold, the production algorithm: which form of its policy runs, on which ring and with how many persistent
// warps is a measured choice (profiles/r02_simple_pipeline_sweep.txt, profiles/r02_halving_slices_probe.txt).
//   * The software-pipelined form (prefetch of the next group's rows + deferred quiet test, PackedSimpleT<true>, 92 registers) on
//     the 2 x 32-tick ring runs every launch since its dependent chain lost an instruction (the accumulator reset as one IMAD):
//     with 20 warps per SM and halving slices it beats the straight-line form with 28 at every link count — 5920 links 73.1 %
//     against 67.8 % of the HBM peak, 6000 links 73.3 / 67.8 %, 8288 links 74.3 / 73.2 %, 4440 links 72.1 / 67.6 % — except on
//     a dense-hit batch (20.2 / 21.0 %).
//   * Below one link per warp of the 20-per-SM grid (2960 links on B200: a 750-link shard of an 8-GPU module, the streaming
//     path's 240 links, one APA) a launch is bound by what ONE warp does per tick and keeps at most 16 warps per SM (measured
//     optimum of the round-2 sweep); beyond that 20 warps per SM and slices even for whole rounds.
// SWTPG_SIMPLE_PIPE (tuning aid): 1 = the pipelined form with the policy's own 16 warps per SM, 0 = the straight-line form on the
// default ring (20 warps per SM), 2 = the straight-line form on the 2 x 16-tick ring with 28 warps per SM (round 2's full-load form).
template<bool DUMP>
cudaError_t
launch_wibeth_simple(const KernelParams& kp, cudaStream_t s)
{
  static const int forced = [] { const char* e = getenv("SWTPG_SIMPLE_PIPE"); return e ? atoi(e) : -1; }();
  if (forced == 0)
    return launch_wibeth_geo<PackedSimpleWibEth, DUMP, GeoDefault>(kp, s);
  if (forced == 1)
    return launch_wibeth_geo<PackedSimpleWibEthPipe, DUMP, GeoDefault>(kp, s);
  if (forced == 2)
    return launch_wibeth_geo<PackedSimpleWibEth, DUMP, Geo<1, 2, 16>, 28>(kp, s);
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess)
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  constexpr int kFullWarpsPerSm = 20;
  if (kp.n_links >= unsigned(kFullWarpsPerSm * sms)) // 2960 links: 66.5 % with 20 warps per SM against 64.2 % with 16; 2500: 56.7 / 58.2 %
    return launch_wibeth_geo<PackedSimpleWibEthPipe, DUMP, GeoDefault, kFullWarpsPerSm, true>(kp, s);
  return launch_wibeth_geo<PackedSimpleWibEthPipe, DUMP, GeoDefault>(kp, s);
}

template<class Algo, bool DUMP>
cudaError_t
launch_wibeth(const KernelParams& kp, cudaStream_t s)
{
  if constexpr (std::is_same<Algo, PackedSimpleWibEth>::value && Algo::kQuadCtasPerSm == 0)
    return launch_wibeth_simple<DUMP>(kp, s);
  if constexpr (Algo::kQuadCtasPerSm > 0) {
    static const bool warp_form = [] {
      const char* e = getenv("SWTPG_WIBETH_KERNEL");
      return e && std::string(e) == "warp";
    }();
    if (!warp_form)
      return launch_wibeth_quad<Algo, DUMP>(kp, s);
  }
  // The running sums (AbsRS, StandardRS; one-warp-per-CTA form): a launch of more than one round of a 20-per-SM grid runs 20
  // warps per SM with halving slices, whole rounds included — AbsRS 28.7 -> 30.4 % of the HBM peak at 5920 links (its own
  // optimum of 16 warps per SM holds for whole links: 20 unsliced 26.9 %), StandardRS 35.4 -> 36.7 %
  // (profiles/r02_warps_crossover_probe.txt).
  if constexpr (std::is_same<Algo, PackedRsWibEth<false>>::value || std::is_same<Algo, PackedRsWibEth<true>>::value) {
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess)
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (kp.n_links > unsigned(20 * sms))
      return launch_wibeth_geo<Algo, DUMP, GeoDefault, 20, true>(kp, s);
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
  kp.link_done = h->d_link_cursor + 2;
  kp.parts_log2 = 0;
  kp.sink.buf = d_tps;
  kp.sink.count = d_count;
  kp.sink.cap = h->tp_capacity;
  kp.pedestal_out = ped;
  kp.waveform_out = wav;
  kp.threshold = h->cfg.threshold;
  kp.acc_limit = h->cfg.frugal_acc_limit;
  kp.acc_limit_neg = 0u - (uint32_t(h->cfg.frugal_acc_limit) & 0xFFFFu);
  kp.rs_scale = int16_t(h->cfg.rs_scale_factor);
  kp.tap_exponent = h->cfg.tap_exponent;
  for (int i = 0; i < 8; ++i)
    kp.taps[i] = h->cfg.fir_taps[i];
  kp.wib2_adc_offset = h->cfg.wib2_adc_offset;
  static const bool force_exact = [] { const char* e = getenv("SWTPG_FIR_FORCE_EXACT"); return e && atoi(e) != 0; }();
  kp.debug_flags = (force_exact || h->fir_force_exact) ? 1u : 0u;
  kp.all_ones = 0xFFFFFFFFu;
  kp.one = 1u;
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
  SW_CUDA(h, cudaMemsetAsync(h->d_link_cursor, 0, (2 + size_t(h->cfg.n_links)) * sizeof(uint32_t), h->stream));
  SW_CUDA(h, cudaStreamSynchronize(h->stream));
  return SWTPG_OK;
}

// Rewrites the SV_RS_FACTOR rows of links [link0, link0 + n) from h->h_rs_factor (or the configured factor): ONE strided
// asynchronous copy on the compute stream, i.e. ordered before every kernel launched after this call and after every kernel
// launched before it. The source is pageable, so the runtime has staged it when the call returns.
swtpg_status
upload_rs_factor_rows(swtpg_handle* h, uint32_t link0, uint32_t n)
{
  const uint32_t g0 = link0 * h->groups_per_link, ng = n * h->groups_per_link;
  std::vector<uint32_t> rows(size_t(ng) * 32);
  for (uint32_t g = 0; g < ng; ++g) {
    const uint32_t link = (g0 + g) / h->groups_per_link, sub = (g0 + g) % h->groups_per_link;
    for (uint32_t lane = 0; lane < 32; ++lane) {
      uint32_t lo = h->cfg.rs_memory_factor, hi = lo;
      if (h->h_rs_factor) {
        const uint16_t* f = h->h_rs_factor + size_t(link) * h->channels + sub * 64 + 2 * lane;
        lo = f[0];
        hi = f[1];
      }
      rows[size_t(g) * 32 + lane] = lo | (hi << 16);
    }
  }
  SW_CUDA(h, cudaMemcpy2DAsync(h->d_state + size_t(g0) * kStateWordsPerGroup + SV_RS_FACTOR * 32, size_t(kStateWordsPerGroup) * 4, rows.data(), 128,
                               128, ng, cudaMemcpyHostToDevice, h->stream));
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
    // pinned source, double-buffered: the copy that last read this buffer (two calls ago) must have finished
    const uint32_t turn = h->nunits_turn++ & 1u;
    SW_CUDA(h, cudaEventSynchronize(h->ev_nunits[turn]));
    memcpy(h->h_nunits[turn], n_units, size_t(h->cfg.n_links) * 4);
    SW_CUDA(h, cudaMemcpyAsync(h->d_nunits, h->h_nunits[turn], size_t(h->cfg.n_links) * 4, cudaMemcpyHostToDevice, s));
    SW_CUDA(h, cudaEventRecord(h->ev_nunits[turn], s));
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
  const swtpg_tp* src = h->d_tps;
  bool finish_on_host = false;
  std::unique_lock<std::mutex> sorter_lock; // the ordered list lives in the sorter's buffer until our copy has completed
  if (h->sorter && n && out) {
    const swtpg_status st = swtpg_internal::sort_tps_device(h, h->sorter, h->d_tps, stored, s, &src, &finish_on_host, &sorter_lock);
    if (st != SWTPG_OK)
      return st;
  }
  if (n && out)
    SW_CUDA(h, cudaMemcpyAsync(out, src, n * sizeof(swtpg_tp), cudaMemcpyDeviceToHost, s));
  SW_CUDA(h, cudaStreamSynchronize(s));
  if (sorter_lock.owns_lock())
    sorter_lock.unlock();
  if (finish_on_host && n == stored)
    swtpg_sort_tps(out, n); // equal keys / keys wider than 64 bits: the host's tie-break makes both orderings identical
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

} // namespace

cudaError_t
swtpg_internal::launch_batch_kernel(swtpg_handle* h, const void* d_frames, const uint32_t* d_nunits, uint32_t units_stride, swtpg_tp* d_tps,
                                    unsigned* d_count, cudaStream_t s)
{
  const KernelParams kp = make_params(h, d_frames, d_nunits, units_stride, d_tps, d_count, nullptr, nullptr);
  return launch<false>(h, kp, s);
}

extern "C" void swtpg_internal_ingest_counts(swtpg_handle* h, uint64_t* zero_copy, uint64_t* staged); // swtpg_stream.cu

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
  if (!h)
    return g_create_error.c_str();
  swtpg_handle* hh = const_cast<swtpg_handle*>(h);
  std::lock_guard<std::mutex> lk(hh->err_mu);
  g_error_copy = hh->last_error; // per-thread copy: another thread may fail (and rewrite the text) at any time
  return g_error_copy.c_str();
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
  if (cfg->flags & ~SWTPG_FLAG_SORTED_TPS)
    return fail(nullptr, SWTPG_ERR_INVALID_ARG, "unknown bits in swtpg_config.flags");

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
    h->cfg.n_slots = 4; // measured: 2 slots 38.0, 3 slots 45.4, 4 slots 47.1, 6 slots 47.2 GB/s (profiles/r02_streaming_slots_probe.txt)
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
    // thresholds sigma * threshold are compared by the bf16 comparator (valid up to 32640) and built by a packed multiply-add
    // (no carry between the halves while (sigmaMax + 3) * threshold < 2^16); outside that range every group takes the
    // exact-threshold tier (plain signed compares on the 64-bit-lane product, like the reference's _mm256_cmpgt_epi16)
    const bool packed_ok = e >= 1 && e <= 10 && getenv("SWTPG_FORCE_SCALAR") == nullptr;
    const bool range_ok = (sigma_max + 3) * cfg->threshold < 65536 && sigma_max * cfg->threshold <= 32640;
    h->fast_rs_wib2 = packed_ok;
    h->fir_force_exact = packed_ok && !range_ok;
  }
  // Packed FIR fast path validity (see PackedFirIqr): binomial taps, and (sigmaMax + 3) * multiplier * threshold < 2^16
  {
    static const int16_t kBinomial[8] = { 1, 6, 15, 20, 15, 6, 1, 0 };
    bool taps_ok = true;
    for (int i = 0; i < 8; ++i)
      taps_ok &= h->cfg.fir_taps[i] == kBinomial[i];
    const uint32_t e = h->cfg.tap_exponent, mult = 1u << e;
    const uint64_t sigma_max = (1u << 15) / (mult * 5u);
    // The packed policies build sigma * multiplier * threshold with one multiply-add (no carry between the halves while
    // (sigmaMax + 3) * K < 2^16) and compare through the bf16 comparator, which orders like signed int16 only up to 32640:
    // K = multiplier * threshold with sigmaMax * K <= 32640 (threshold <= 5 at the reference's exponent 6). Beyond that the
    // reference's _mm256_cmpgt_epi16 sees the product as a NEGATIVE int16 once a channel's IQR is large enough, so such
    // configurations run the same packed trackers and filter with the exact-threshold tier forced on for every group.
    const uint64_t K = uint64_t(mult) * cfg->threshold;
    const bool packed_ok = cfg->algorithm == SWTPG_ALGO_FIR_IQR && e >= 1 && e <= 10 && getenv("SWTPG_FORCE_SCALAR") == nullptr;
    const bool range_ok = (sigma_max + 3) * K < 65536 && sigma_max * K <= 32640;
    h->fast_fir = packed_ok && taps_ok;      // binomial cascade
    h->fast_fir_any = packed_ok && !taps_ok; // any other taps[0..6]: packed multiply-add chain (PackedFirIqrAnyTaps)
    if (packed_ok && !range_ok)
      h->fir_force_exact = true;
  }

  swtpg_handle* hp = h.get();
  SW_CUDA(hp, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  SW_CUDA(hp, cudaEventCreate(&h->ev0));
  SW_CUDA(hp, cudaEventCreate(&h->ev1));
  SW_CUDA(hp, cudaMalloc(&h->d_state, size_t(h->n_groups) * kStateWordsPerGroup * 4));
  SW_CUDA(hp, cudaMalloc(&h->d_flags, size_t(h->n_groups) * 4));
  SW_CUDA(hp, cudaMalloc(&h->d_link_cursor, (2 + size_t(h->cfg.n_links)) * sizeof(uint32_t)));
  SW_CUDA(hp, cudaMemset(h->d_link_cursor, 0, (2 + size_t(h->cfg.n_links)) * sizeof(uint32_t)));
  SW_CUDA(hp, cudaMalloc(&h->d_tps, size_t(h->tp_capacity) * sizeof(swtpg_tp)));
  SW_CUDA(hp, cudaMalloc(&h->d_count, sizeof(unsigned)));
  SW_CUDA(hp, cudaMallocHost(&h->h_count, sizeof(unsigned)));
  SW_CUDA(hp, cudaMalloc(&h->d_nunits, size_t(cfg->n_links) * 4));
  for (int i = 0; i < 2; ++i) {
    SW_CUDA(hp, cudaMallocHost(&h->h_nunits[i], size_t(cfg->n_links) * 4));
    SW_CUDA(hp, cudaEventCreateWithFlags(&h->ev_nunits[i], cudaEventDisableTiming));
  }
  if (cfg->flags & SWTPG_FLAG_SORTED_TPS)
    h->sorter = swtpg_internal::sorter_create();
  *out = h.release();
  return SWTPG_OK;
}

void
swtpg_destroy(swtpg_handle* h)
{
  if (!h)
    return;
  cudaSetDevice(h->cfg.device);
  swtpg_internal::engine_destroy(h); // joins the streaming threads, unregisters latency buffers
  cudaDeviceSynchronize();
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
  for (int i = 0; i < 2; ++i) {
    if (h->h_nunits[i]) cudaFreeHost(h->h_nunits[i]);
    if (h->ev_nunits[i]) cudaEventDestroy(h->ev_nunits[i]);
  }
  if (h->d_frames) cudaFree(h->d_frames);
  if (h->d_ped) cudaFree(h->d_ped);
  if (h->d_wav) cudaFree(h->d_wav);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->sorter) swtpg_internal::sorter_destroy(h->sorter);
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
  s = swtpg_internal::engine_reset(h); // streaming path: idle, rings empty, undelivered TPs of the previous run dropped
  if (s != SWTPG_OK)
    return s;
  h->counters.reset();
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
  if (h->started) { // only the factor rows change; carried state is preserved
    SW_CUDA(h, cudaSetDevice(h->cfg.device));
    return upload_rs_factor_rows(h, 0, h->cfg.n_links);
  }
  return SWTPG_OK;
}

swtpg_status
swtpg_set_link_rs_memory_factor(swtpg_handle* h, uint32_t link, const uint16_t* by_channel)
{
  if (!h || !by_channel || link >= h->cfg.n_links)
    return SWTPG_ERR_INVALID_ARG;
  {
    std::lock_guard<std::mutex> lk(h->engine_mu); // several links' threads may arrive here with their first frames
    if (!h->h_rs_factor) {
      const size_t n = size_t(h->cfg.n_links) * h->channels;
      h->h_rs_factor = new uint16_t[n];
      std::fill(h->h_rs_factor, h->h_rs_factor + n, h->cfg.rs_memory_factor);
    }
    memcpy(h->h_rs_factor + size_t(link) * h->channels, by_channel, size_t(h->channels) * sizeof(uint16_t));
  }
  if (!h->started)
    return SWTPG_OK; // swtpg_start uploads the table
  SW_CUDA(h, cudaSetDevice(h->cfg.device));
  return upload_rs_factor_rows(h, link, 1);
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

// Copy of one payload into pinned memory (csrc/swtpg_hostutil.cpp: non-temporal stores where the CPU has AVX2).
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

swtpg_status
swtpg_sync(swtpg_handle* h)
{
  if (!h)
    return SWTPG_ERR_INVALID_ARG;
  SW_CUDA(h, cudaSetDevice(h->cfg.device));
  swtpg_status s = swtpg_internal::engine_quiesce(h); // streaming path: every dispatched batch has delivered its TPs to the ready queue
  if (s != SWTPG_OK)
    return s;
  SW_CUDA(h, cudaStreamSynchronize(h->stream));
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
  out->units_processed = h->counters.units_processed.load();
  out->samples_processed = h->counters.samples_processed.load();
  out->tps_emitted = h->counters.tps_emitted.load();
  out->tps_dropped_overflow = h->counters.tps_dropped_overflow.load();
  out->batches = h->counters.batches.load();
  out->submit_busy = h->counters.submit_busy.load();
  out->h2d_bytes = h->counters.h2d_bytes.load();
  out->d2h_bytes = h->counters.d2h_bytes.load();
  swtpg_internal_ingest_counts(h, &out->units_zero_copy, &out->units_staged);
  return SWTPG_OK;
}

swtpg_status
swtpg_sort_stats(swtpg_handle* h, double* last_ms, double* total_ms, uint64_t* lists, uint64_t* finished_on_host)
{
  if (!h)
    return SWTPG_ERR_INVALID_ARG;
  swtpg_internal::sorter_stats(h->sorter, last_ms, total_ms, lists, finished_on_host);
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

} // extern "C"
