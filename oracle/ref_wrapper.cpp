// TEST INFRASTRUCTURE — not part of the product path.
//
// Thin C-ABI glue that compiles the UNMODIFIED reference SWTPG headers, in place under /root/reference/include,
// into oracle/_ref/libswtpg_ref.so (see oracle/Makefile). Nothing here restates the reference's arithmetic: the
// hit finders are the reference's own templates; only the per-frame driver (what WIBEthFrameProcessor::find_hits /
// WIB2FrameProcessor::find_hits do around them) and the hit-block decode (process_swtpg_hits) are restated,
// because src/**/…FrameProcessor.cpp cannot be built here (readoutlibs, iomanager, appfwk, ers are absent).
//
//   driver  : src/wibeth/WIBEthFrameProcessor.cpp:410-476, src/wib2/WIB2FrameProcessor.cpp:345-396
//   decode  : src/wibeth/WIBEthFrameProcessor.cpp:478-572, src/wib2/WIB2FrameProcessor.cpp:398-479
//   handler : src/wibeth/WIBEthFrameProcessor.cpp:74-91 (100000-word hit buffer, exponent 6),
//             src/wib2/WIB2FrameProcessor.cpp:90-120 (taps = firwin_int(7, 0.1, 64) + {0})
//
// Used by tests/ (to pin oracle/swtpg_oracle.c and as a second checker for the CUDA path) and by
// bench.py's cpu_baseline / --impl reference legs. Never linked or loaded by fdreadoutlibs_b200's product code.

#include "fdreadoutlibs/wibeth/tpg/FrameExpand.hpp"
#include "fdreadoutlibs/wibeth/tpg/ProcessingInfo.hpp"
#include "fdreadoutlibs/wibeth/tpg/ProcessAVX2.hpp"
#include "fdreadoutlibs/wibeth/tpg/ProcessAbsRSAVX2.hpp"
#include "fdreadoutlibs/wibeth/tpg/ProcessStandardRSAVX2.hpp"
#include "fdreadoutlibs/wibeth/tpg/ProcessNaive.hpp" // defines a non-inline function: this TU only

#include "fdreadoutlibs/wib2/tpg/FrameExpand.hpp"
#include "fdreadoutlibs/wib2/tpg/ProcessingInfo.hpp"
#include "fdreadoutlibs/wib2/tpg/ProcessAVX2.hpp"
#include "fdreadoutlibs/wib2/tpg/ProcessAVX2FIR.hpp"
#include "fdreadoutlibs/wib2/tpg/ProcessRSAVX2.hpp"
#include "fdreadoutlibs/wib2/tpg/ProcessNaive.hpp" // ditto
#include <cmath>   // ProcessNaiveRS.hpp uses std::round and std::stringstream without including their headers
#include <sstream>
// The reference's scalar AbsRS processor for WIB2. NOT a scalar twin of wib2/tpg/ProcessRSAVX2.hpp: it keeps the running sum
// in float (R = 0.8, scale 2), tracks the inter-quartile range of the RUNNING SUM rather than of the raw samples, hard-wires the
// threshold 5 * sigma and accumulates the pedestal-subtracted sample (not the running sum) as charge, so its hits differ from
// the AVX2 processor's by construction (tests/test_oracle_vs_reference.py quantifies it). Nothing is pinned to it.
#include "fdreadoutlibs/wib2/tpg/ProcessNaiveRS.hpp"
#include "fdreadoutlibs/wib2/tpg/DesignFIR.hpp"

#include "../include/swtpg.h"

#include <pthread.h>
#include <sched.h>

#include <array>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace {

// Lane l of register r holds frame channel 16r + kPerm[l] (unittest/WIBEthFrameExpansion_test.cxx:111,124).
constexpr int kPerm[16] = { 0, 1, 2, 3, 4, 5, 6, 7, 15, 8, 9, 10, 11, 12, 13, 14 };
inline int
position_to_channel(int pos)
{
  return (pos & ~15) | kPerm[pos & 15];
}

// setState writes a line to std::cout per link (wibeth/tpg/ProcessingInfo.hpp:110,146). bench.py must print exactly one
// JSON line on stdout, so this library points std::cout at a null buffer once, when the first handle is created.
struct NullBuf : std::streambuf
{
  int overflow(int c) override { return c; }
};
void
silence_cout()
{
  static std::once_flag once;
  static NullBuf nullbuf;
  std::call_once(once, [] { std::cout.rdbuf(&nullbuf); });
}

enum WibEthImpl
{
  kEthSimpleAVX2 = 0,
  kEthSimpleNaive = 1,
  kEthAbsRSAVX2 = 2,
  kEthStandardRSAVX2 = 3
};

struct WibEthRef
{
  int impl;
  bool first_hit = true;
  std::unique_ptr<uint16_t[]> hits;
  std::unique_ptr<swtpg_wibeth::ProcessingInfo<swtpg_wibeth::NUM_REGISTERS_PER_FRAME>> info;
  std::array<uint16_t, 64> memory_factor; // per register position, src/wibeth/WIBEthFrameProcessor.cpp:437-456
  uint64_t n_blocks = 0;
};

enum Wib2Impl
{
  kWib2SimpleAVX2 = 0,
  kWib2FirAVX2 = 1,
  kWib2FirNaive = 2,
  kWib2AbsRSAVX2 = 3,
  kWib2AbsRSNaive = 4 // wib2/tpg/ProcessNaiveRS.hpp: float running sum, IQR of the running sum, threshold 5 * sigma
};

struct Wib2Ref
{
  int impl;
  int sel;
  bool first_hit = true;
  std::vector<int16_t> taps;
  std::unique_ptr<uint16_t[]> hits;
  std::unique_ptr<swtpg_wib2::ProcessingInfo<swtpg_wib2::NUM_REGISTERS_PER_FRAME>> info;
};

// Decode of 7-field WIBEth AVX2 hit blocks (src/wibeth/WIBEthFrameProcessor.cpp:487-549). `chan` is already the
// frame channel (wibeth/tpg/ProcessAVX2.hpp:32,67-68). Returns number of TPs appended, or -1 on overflow of `cap`.
long
decode_wibeth_avx2(const uint16_t* it, uint64_t timestamp, uint32_t link, swtpg_tp* out, size_t cap, size_t n)
{
  const size_t n0 = n;
  while (*it != swtpg_wibeth::MAGIC) {
    const uint16_t* chan = it;
    const uint16_t* hit_end = it + 16;
    const uint16_t* hit_charge = it + 32;
    const uint16_t* hit_tover = it + 48;
    const uint16_t* hit_peak_adc = it + 64;
    const uint16_t* hit_peak_time = it + 80;
    const uint16_t* left = it + 96;
    it += 112;
    for (int i = 0; i < 16; ++i) {
      if (hit_charge[i] && left[i] == swtpg_wibeth::MAGIC && chan[i] != swtpg_wibeth::MAGIC) {
        if (n >= cap)
          return -1;
        swtpg_tp& tp = out[n++];
        tp.time_start = timestamp + 32 * ((int64_t)hit_end[i] - (int64_t)hit_tover[i]);
        tp.time_peak = tp.time_start + 32 * hit_peak_time[i];
        tp.time_over_threshold = uint32_t(hit_tover[i]) * 32;
        tp.adc_integral = hit_charge[i];
        tp.adc_peak = hit_peak_adc[i];
        tp.channel = chan[i];
        tp.link = link;
      }
    }
  }
  return long(n - n0);
}

// Decode of the 6-field scalar records of process_window_naive (wibeth/tpg/ProcessNaive.hpp:113-118); the record's
// first word is a register POSITION, mapped here to the frame channel. Same TP derivation as above.
long
decode_wibeth_naive(const uint16_t* it, uint64_t timestamp, uint32_t link, swtpg_tp* out, size_t cap, size_t n)
{
  const size_t n0 = n;
  while (*it != swtpg_wibeth::MAGIC) {
    const uint16_t pos = it[0], hit_end = it[1], charge = it[2], tover = it[3], peak_adc = it[4], peak_time = it[5];
    it += 6;
    if (!charge)
      continue;
    if (n >= cap)
      return -1;
    swtpg_tp& tp = out[n++];
    tp.time_start = timestamp + 32 * ((int64_t)hit_end - (int64_t)tover);
    tp.time_peak = tp.time_start + 32 * peak_time;
    tp.time_over_threshold = uint32_t(tover) * 32;
    tp.adc_integral = charge;
    tp.adc_peak = peak_adc;
    tp.channel = uint16_t(position_to_channel(pos));
    tp.link = link;
  }
  return long(n - n0);
}

// WIB2 4-field blocks (src/wib2/WIB2FrameProcessor.cpp:408-458). The first word of a lane is
// position + chan_bias, where chan_bias = 128*sel for the SimpleThreshold / AbsRS AVX2 kernels (sequential iota plus
// channel_offset, wib2/tpg/ProcessAVX2.hpp:33,64) and 0 for the FIR kernels (wib2/tpg/ProcessAVX2FIR.hpp:91-92,
// wib2/tpg/ProcessNaive.hpp:140). The production LUT is position-indexed (src/wib2/WIB2FrameProcessor.cpp:367-368),
// so the frame channel is perm(position) + 128*sel.
long
decode_wib2(const uint16_t* it, bool scalar_records, int chan_bias, int sel, uint64_t timestamp, uint32_t link, swtpg_tp* out, size_t cap, size_t n)
{
  const size_t n0 = n;
  auto emit = [&](uint16_t chan, uint16_t hit_end, uint16_t charge, uint16_t tover) -> bool {
    if (n >= cap)
      return false;
    swtpg_tp& tp = out[n++];
    const uint64_t t_begin = timestamp + 32 * (int64_t(hit_end) - int64_t(tover));
    const uint64_t t_end = timestamp + 32 * int64_t(hit_end);
    tp.time_start = t_begin;
    tp.time_peak = (t_begin + t_end) / 2;
    tp.time_over_threshold = uint32_t(int64_t(tover) * 32);
    tp.adc_integral = charge;
    tp.adc_peak = charge / 20;
    tp.channel = uint16_t(position_to_channel(chan - chan_bias) + 128 * sel);
    tp.link = link;
    return true;
  };
  while (*it != swtpg_wib2::MAGIC) {
    if (scalar_records) { // wib2/tpg/ProcessNaive.hpp:140-143: {position, itime, charge, tover}
      if (it[2] && !emit(it[0], it[1], it[2], it[3]))
        return -1;
      it += 4;
    } else {
      for (int i = 0; i < 16; ++i) {
        if (it[32 + i] && it[i] != swtpg_wib2::MAGIC) {
          if (!emit(it[i], it[16 + i], it[32 + i], it[48 + i]))
            return -1;
        }
      }
      it += 64;
    }
  }
  return long(n - n0);
}

inline void
process_one_wibeth_frame(WibEthRef* r, const uint8_t* frame)
{
  using namespace swtpg_wibeth;
  auto fp = reinterpret_cast<const dunedaq::fdreadoutlibs::types::DUNEWIBEthTypeAdapter*>(frame);
  MessageRegisters registers_array;
  expand_wibeth_adcs(fp, &registers_array);
  if (r->first_hit) { // src/wibeth/WIBEthFrameProcessor.cpp:424-464
    r->info->setState(registers_array, r->memory_factor);
    r->first_hit = false;
  }
  r->info->input = &registers_array;
  r->info->output[0] = MAGIC;
  switch (r->impl) {
    case kEthSimpleAVX2: process_window_avx2(*r->info); break;
    case kEthSimpleNaive: process_window_naive(*r->info); break;
    case kEthAbsRSAVX2: process_window_rs_avx2(*r->info); break;
    default: process_window_standard_rs_avx2(*r->info); break;
  }
  r->n_blocks += r->info->nhits;
}

} // namespace

extern "C" {

void*
ref_wibeth_create(int impl, uint16_t threshold, int16_t acc_limit, uint16_t memory_factor, uint16_t scale_factor)
{
  silence_cout();
  auto* r = new WibEthRef;
  r->impl = impl;
  r->hits.reset(new uint16_t[100000]); // src/wibeth/WIBEthFrameProcessor.cpp:78
  r->memory_factor.fill(memory_factor);
  r->info = std::make_unique<swtpg_wibeth::ProcessingInfo<swtpg_wibeth::NUM_REGISTERS_PER_FRAME>>(
    nullptr, swtpg_wibeth::FRAMES_PER_MSG, 0, swtpg_wibeth::NUM_REGISTERS_PER_FRAME, r->hits.get(), 6, threshold,
    memory_factor, scale_factor, acc_limit, 0);
  return r;
}

void
ref_wibeth_destroy(void* h)
{
  delete static_cast<WibEthRef*>(h);
}

// Per-position memory factor (collection-plane channels get 0 when enable_simple_threshold_on_collection,
// src/wibeth/WIBEthFrameProcessor.cpp:441-450). `by_channel` is indexed by FRAME channel.
void
ref_wibeth_set_memory_factor(void* h, const uint16_t* by_channel)
{
  auto* r = static_cast<WibEthRef*>(h);
  for (int pos = 0; pos < 64; ++pos)
    r->memory_factor[pos] = by_channel[position_to_channel(pos)];
}

// Run n_frames consecutive 7200-byte frames of one link. TPs are appended to out[0..cap). If ped_dump != NULL it
// receives, after every frame, the 64 pedestals and 64 accumulators in FRAME-CHANNEL order:
// ped_dump[(f*2+0)*64 + ch] = pedestal, ped_dump[(f*2+1)*64 + ch] = accum.
long
ref_wibeth_process(void* h, const uint8_t* frames, size_t n_frames, uint32_t link, swtpg_tp* out, size_t cap, int16_t* ped_dump)
{
  auto* r = static_cast<WibEthRef*>(h);
  size_t n = 0;
  for (size_t f = 0; f < n_frames; ++f) {
    const uint8_t* frame = frames + f * 7200;
    uint64_t ts;
    memcpy(&ts, frame + 8, 8);
    process_one_wibeth_frame(r, frame);
    long k = (r->impl == kEthSimpleNaive) ? decode_wibeth_naive(r->info->output, ts, link, out, cap, n)
                                          : decode_wibeth_avx2(r->info->output, ts, link, out, cap, n);
    if (k < 0)
      return -1;
    n += size_t(k);
    if (ped_dump) {
      for (int pos = 0; pos < 64; ++pos) {
        const int ch = position_to_channel(pos);
        ped_dump[(f * 2 + 0) * 64 + ch] = r->info->chanState.pedestals[pos];
        ped_dump[(f * 2 + 1) * 64 + ch] = r->info->chanState.accum[pos];
      }
    }
  }
  return long(n);
}

// Raw expansion of one frame: out[4096] = MessageRegisters contents (wibeth/tpg/FrameExpand.hpp:192-246).
void
ref_wibeth_expand(const uint8_t* frame, uint16_t* out)
{
  swtpg_wibeth::MessageRegisters regs;
  swtpg_wibeth::expand_wibeth_adcs(reinterpret_cast<const dunedaq::fdreadoutlibs::types::DUNEWIBEthTypeAdapter*>(frame), &regs);
  memcpy(out, regs.data(), 4096 * sizeof(uint16_t));
}

void*
ref_wib2_create(int impl, uint16_t threshold, int sel)
{
  silence_cout();
  auto* r = new Wib2Ref;
  r->impl = impl;
  r->sel = sel;
  r->taps = swtpg_wib2::firwin_int(7, 0.1, 64); // src/wib2/WIB2FrameProcessor.cpp:93-94, multiplier 1<<6
  r->taps.push_back(0);
  r->hits.reset(new uint16_t[100000]);
  r->info = std::make_unique<swtpg_wib2::ProcessingInfo<swtpg_wib2::NUM_REGISTERS_PER_FRAME>>(
    nullptr, swtpg_wib2::FRAMES_PER_MSG, 0, swtpg_wib2::NUM_REGISTERS_PER_FRAME, r->hits.get(), r->taps.data(),
    (uint8_t)r->taps.size(), 6, threshold, 0, 0);
  return r;
}

// Same with caller-supplied taps and exponent (the reference's ProcessingInfo takes both; its frame processor just never
// passes anything but firwin_int(7, 0.1, 64) and 6): pins the oracle's filter arithmetic for arbitrary taps.
void*
ref_wib2_create_taps(int impl, uint16_t threshold, int sel, const int16_t* taps8, int tap_exponent)
{
  silence_cout();
  auto* r = new Wib2Ref;
  r->impl = impl;
  r->sel = sel;
  r->taps.assign(taps8, taps8 + 8);
  r->hits.reset(new uint16_t[100000]);
  r->info = std::make_unique<swtpg_wib2::ProcessingInfo<swtpg_wib2::NUM_REGISTERS_PER_FRAME>>(
    nullptr, swtpg_wib2::FRAMES_PER_MSG, 0, swtpg_wib2::NUM_REGISTERS_PER_FRAME, r->hits.get(), r->taps.data(),
    (uint8_t)r->taps.size(), (uint8_t)tap_exponent, threshold, 0, 0);
  return r;
}

void
ref_wib2_destroy(void* h)
{
  delete static_cast<Wib2Ref*>(h);
}

// Run n_sc consecutive 5664-byte superchunks (12 WIB2 frames) through handler `sel` (channels 128*sel..+127).
// state_dump (optional): after every superchunk, pedestal[128] then (FIR only) q25[128], q75[128] in channel order
// relative to 128*sel: state_dump[(s*3+k)*128 + c].
long
ref_wib2_process(void* h, const uint8_t* superchunks, size_t n_sc, uint32_t link, swtpg_tp* out, size_t cap, int16_t* state_dump)
{
  using namespace swtpg_wib2;
  auto* r = static_cast<Wib2Ref*>(h);
  size_t n = 0;
  for (size_t s = 0; s < n_sc; ++s) {
    const uint8_t* sc = superchunks + s * 5664;
    auto fp = reinterpret_cast<const dunedaq::fdreadoutlibs::types::DUNEWIBSuperChunkTypeAdapter*>(sc);
    const uint64_t ts = reinterpret_cast<const dunedaq::fddetdataformats::WIB2Frame*>(sc)->get_timestamp();
    MessageRegisters registers_array;
    expand_wib2_adcs(fp, &registers_array, r->sel);
    if (r->first_hit) {
      r->info->setState(registers_array);
      r->first_hit = false;
    }
    r->info->input = &registers_array;
    r->info->output[0] = MAGIC;
    const size_t off = size_t(r->sel) * NUM_REGISTERS_PER_FRAME * SAMPLES_PER_REGISTER;
    switch (r->impl) {
      case kWib2SimpleAVX2: process_window_avx2(*r->info, off); break;
      case kWib2FirAVX2: process_window_avx2(*r->info); break;
      case kWib2FirNaive: process_window_naive(*r->info, off); break;
      case kWib2AbsRSNaive: process_window_naive_RS(*r->info, off); break; // prints one "Found N hits" line per call (:224)
      default: process_window_rs_avx2(*r->info, off); break;
    }
    const bool scalar = (r->impl == kWib2FirNaive || r->impl == kWib2AbsRSNaive); // {position, itime, charge, tover} records
    const bool positions = scalar || r->impl == kWib2FirAVX2;                       // channel ids relative to the handler's 128
    long k = decode_wib2(r->info->output, scalar, positions ? 0 : 128 * r->sel, r->sel, ts, link, out, cap, n);
    if (k < 0)
      return -1;
    n += size_t(k);
    if (state_dump) {
      for (int pos = 0; pos < 128; ++pos) {
        const int ch = position_to_channel(pos);
        state_dump[(s * 3 + 0) * 128 + ch] = r->info->chanState.pedestals[pos];
        state_dump[(s * 3 + 1) * 128 + ch] = r->info->chanState.quantile25[pos];
        state_dump[(s * 3 + 2) * 128 + ch] = r->info->chanState.quantile75[pos];
      }
    }
  }
  return long(n);
}

void
ref_wib2_expand(const uint8_t* superchunk, int sel, uint16_t* out /* 8*12*16 */)
{
  swtpg_wib2::MessageRegisters regs;
  swtpg_wib2::expand_wib2_adcs(reinterpret_cast<const dunedaq::fdreadoutlibs::types::DUNEWIBSuperChunkTypeAdapter*>(superchunk), &regs, sel);
  memcpy(out, regs.data(), 8 * 12 * 16 * sizeof(uint16_t));
}

int
ref_firwin_int(int n, double cutoff, int multiplier, int16_t* out)
{
  auto t = swtpg_wib2::firwin_int(n, cutoff, multiplier);
  for (size_t i = 0; i < t.size(); ++i)
    out[i] = t[i];
  return int(t.size());
}

// CPU baseline: the reference's own per-link loop (expand + process_window + hit decode), one worker thread per
// group of links, pinned, mirroring "one post-processing thread per link" (src/wibeth/WIBEthFrameProcessor.cpp:231).
// frames = [n_links][n_frames][7200]. Each of `reps` passes re-creates the per-link state (like start()).
// Returns the best wall-clock seconds over reps; *n_tps gets the TP count of one pass.
double
ref_wibeth_bench(const uint8_t* frames, size_t n_links, size_t n_frames, int n_threads, int impl, uint16_t threshold,
                 int16_t acc_limit, int reps, uint64_t* n_tps)
{
  if (n_threads < 1)
    n_threads = 1;
  double best = 1e30;
  for (int rep = 0; rep < reps; ++rep) {
    std::vector<void*> handles(n_links);
    for (auto& hd : handles)
      hd = ref_wibeth_create(impl, threshold, acc_limit, 0, 0);
    std::atomic<uint64_t> total{ 0 };
    std::atomic<int> ready{ 0 };
    std::atomic<bool> go{ false };
    std::vector<std::thread> workers;
    for (int t = 0; t < n_threads; ++t) {
      workers.emplace_back([&, t]() {
        cpu_set_t set;
        CPU_ZERO(&set);
        CPU_SET(t % std::max(1u, std::thread::hardware_concurrency()), &set);
        pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
        std::vector<swtpg_tp> tps(1 << 16);
        ready.fetch_add(1);
        while (!go.load(std::memory_order_acquire)) {
        }
        uint64_t mine = 0;
        for (size_t l = t; l < n_links; l += n_threads) {
          const uint8_t* base = frames + l * n_frames * 7200;
          for (size_t f = 0; f < n_frames; ++f) { // frame at a time: bounded TP buffer, like the reference
            long k = ref_wibeth_process(handles[l], base + f * 7200, 1, uint32_t(l), tps.data(), tps.size(), nullptr);
            if (k > 0)
              mine += uint64_t(k);
          }
        }
        total.fetch_add(mine);
      });
    }
    while (ready.load() < n_threads) {
    }
    auto t0 = std::chrono::steady_clock::now();
    go.store(true, std::memory_order_release);
    for (auto& w : workers)
      w.join();
    auto t1 = std::chrono::steady_clock::now();
    best = std::min(best, std::chrono::duration<double>(t1 - t0).count());
    if (n_tps)
      *n_tps = total.load();
    for (auto hd : handles)
      ref_wibeth_destroy(hd);
  }
  return best;
}

} // extern "C"
