// Fused SWTPG kernels for sm_100a: 14-bit unpack -> frugal pedestal -> [RS | FIR] -> threshold hit finding -> TP list.
//
// One warp owns one 64-channel group for the whole batch (time is a sequential recurrence per channel; parallelism
// comes from channels and links only). Frame bytes reach the warp through a private ring of shared-
// memory stages filled by the TMA engine (cp.async.bulk + mbarrier complete_tx), so no thread ever waits on a
// global load inside the tick loop and no block-wide barrier exists.
//
// Algorithms are policy structs with the same shape:
//   struct Algo { load(state, lane); seed(S); tick<...>(S, ctx, t); store(state, lane); }
// `Scalar*` policies do the arithmetic one channel at a time in 32-bit registers, exactly as written in the
// reference (any configuration); `Packed*` policies are the production fast paths on packed s16x2 registers and are
// selected by the host only for configurations inside their documented validity range (swtpg_capi.cu: pick_kernel).
#pragma once

#include "swtpg_device.cuh"

#include <type_traits>

#ifndef SWTPG_STEP_NEGL
#define SWTPG_STEP_NEGL 1 // integer frugal step of SimpleThreshold / the running sums: accumulator reset as ONE IMAD with -L (0: c - L * D)
#endif
#ifndef SWTPG_SLICE_FENCES
#define SWTPG_SLICE_FENCES 2 // sliced hand-out: 2 = every lane fences before lane 0's release store; 1 = the release alone
#endif
#ifndef SWTPG_GROUP_UNROLL
#define SWTPG_GROUP_UNROLL 4
#endif
#ifndef SWTPG_FIR_GROUP_UNROLL
#define SWTPG_FIR_GROUP_UNROLL 2
#endif
#ifndef SWTPG_QUAD_SIMPLE
#define SWTPG_QUAD_SIMPLE 0
#endif
#ifndef SWTPG_ELECT
#define SWTPG_ELECT 1
#endif
#ifndef SWTPG_RS_GROUP_UNROLL
#define SWTPG_RS_GROUP_UNROLL 2
#endif
#ifndef SWTPG_FLOAT_ACC
#define SWTPG_FLOAT_ACC 1
#endif
// Frugal accumulator of the SimpleThreshold / running-sum pedestal: 0 = fp16x2 subnormal (round 1), 1 = biased integer, seven
// instructions per tick, 2 = biased integer with the step taken off the median's dependent chain (eight instructions, shorter chain)
#ifndef SWTPG_SIMPLE_INT
#define SWTPG_SIMPLE_INT 1
#endif

namespace swtpg {

// ---- HBM-resident per-group state (struct of arrays: [group][var][lane], u32 = two packed channels) -------------
enum StateVar
{
  SV_MEDIAN = 0,
  SV_ACCUM,
  SV_PREV,
  SV_CHARGE,
  SV_TOVER,
  SV_PEAK_ADC,
  SV_PEAK_TIME,
  SV_Q25,
  SV_Q75,
  SV_A25,
  SV_A75,
  SV_RS,
  SV_MED_RS,
  SV_ACC_RS,
  SV_RS_FACTOR,
  SV_RING0, // .. SV_RING0 + 7
  SV_COUNT = SV_RING0 + 8
};
constexpr uint32_t kStateWordsPerGroup = SV_COUNT * 32;
constexpr uint32_t kFlagInitialized = 1u;

struct KernelParams
{
  const uint8_t* frames;   // link-major units
  const uint32_t* n_units; // per link, or nullptr = units_stride everywhere
  uint32_t units_stride;
  uint32_t n_links;
  uint32_t* state;         // [n_groups][SV_COUNT][32]
  uint32_t* group_flags;   // [n_groups]: bit 0 initialized, bits 8..10 FIR ring phase (absTimeModNTAPS)
  uint32_t* link_cursor;   // [2] {links claimed beyond each warp's first, warps finished}; zero between launches (WIBEth kernel)
  uint32_t* link_done;     // [n_links] slices of the link finished in THIS launch; zero between launches (wibeth_kernel, sliced)
  uint32_t parts_log2;     // wibeth_kernel: a link's units are handed out in 2^parts_log2 consecutive slices (0 = whole links)
  uint32_t slice_geom;     // ... of equal length (0) or halving: n/2, n/4, ..., and the rest (1)
  TpSink sink;
  int16_t* pedestal_out;   // debug dumps [link][unit][tick][channel] or nullptr
  int16_t* waveform_out;
  // configuration (swtpg_config)
  uint32_t threshold;      // u16
  int32_t acc_limit;       // i16
  uint32_t acc_limit_neg;  // -acc_limit as its own parameter: the integer tracker's IMAD then takes it straight from the constant
                           // bank (derived from acc_limit in the kernel it costs a register or a negation per group of ticks)
  int32_t rs_scale;        // i16
  int32_t tap_exponent;
  int32_t taps[8];
  uint32_t wib2_adc_offset;
  uint32_t debug_flags;    // bit 0: PackedFirIqr always takes its exact-threshold tier (test aid, SWTPG_FIR_FORCE_EXACT=1)
  uint32_t all_ones;       // 0xFFFFFFFF: a constant the compiler cannot fold, so that it stays in ONE register (VIADDMNMX takes a
                           // single immediate; ptxas otherwise re-materialises the -1 operand with a move in front of every use)
  uint32_t one;            // 1: multiplier of the adds that are to run on the FMA pipe as IMAD (see fma_add below)
};

// Pipe steering. Measured with ncu on every kernel of this file: the ALU pipe (LOP3, SHF, PRMT, IADD3, VIMNMX, VIADDMNMX, HSET2,
// ISETP, SEL: half issue rate) is 70-75 % busy while the FMA pipe (IMAD, VIADD.16x2, HFMA2, HADD2: half rate as well) idles at
// 25-30 %, and ptxas turns `a + b` into IADD3 (ALU) or IMAD.IADD (FMA) by its own alternation. An add written as a * ONE + b with
// ONE = 1 read from the kernel parameters cannot be strength-reduced and is an IMAD: the add runs on the idle pipe.
// Used by the FIR + IQR trackers (WIB2 layout 30.5 -> 31.8 % of the HBM peak); the SimpleThreshold step keeps plain adds (steered:
// 1 % slower, profiles/r02_integer_accumulators.txt).
__device__ __forceinline__ uint32_t
fma_add(uint32_t a, uint32_t b, uint32_t one)
{
  return a * one + b;
}
__device__ __forceinline__ uint32_t
fma_sub(uint32_t a, uint32_t b, uint32_t minus_one)
{ // a - b
  return b * minus_one + a;
}

struct TickCtx
{
  uint64_t ts;      // timestamp of the current unit
  uint32_t tick_base; // ticks of this batch before the current unit (FIR ring phase)
  uint32_t link;
  uint32_t chan0;   // frame channel of this lane's low half
  uint32_t unit;    // index of the current unit inside the batch
  const uint8_t* link_base;
  HitStage* stage;  // warp-private hit staging (packed policies)
  const KernelParams* p;
};

// =====================================================================================================================
// Scalar policies: the reference arithmetic, one channel at a time. Used for configurations outside the packed fast
// paths' validity range and as the in-kernel statement of the semantics (cf. oracle/swtpg_oracle.c).
// =====================================================================================================================

// wibeth/tpg/UtilsAVX2.hpp:24-74 for one lane; mask = lane participates
__device__ __forceinline__ void
frugal_scalar(int& median, int s, int& accum, int L, bool mask)
{
  int to_add = s > median ? 1 : (s == median ? 0 : -1);
  if (!mask)
    to_add = 0;
  accum = wrap16(accum + to_add);
  const bool is_gt = accum > L;
  const int b = wrap16(-L); // _mm256_set1_epi16(-1 * acclimit)
  const int sa = b < 0 ? wrap16(-accum) : (b == 0 ? 0 : accum); // _mm256_sign_epi16
  const bool is_lt = sa > L;
  int step = is_gt ? 1 : 0;
  if (is_lt)
    step = -1;
  if (!mask)
    step = 0;
  median = sat16(median + step);
  if ((is_gt || is_lt) && mask)
    accum = 0;
}

// Threshold of the FIR + IQR finder for the two channels of a lane: the 16-bit lanes of `sigma * multiplier * threshold`
// evaluated the way GCC evaluates `__m256i * int` — as 4 x int64 lanes (wib2/tpg/ProcessAVX2FIR.hpp:208): the
// sigmas of 4 adjacent register POSITIONS form one u64 that is multiplied mod 2^64, so carries run from one position into
// the next. Position q of a register holds channel perm[q] of its 16 (perm = {0..7,15,8..14}). Whole warp calls;
// sig_packed = this lane's two sigmas (s16x2), lane = index inside the 64-channel group.
__device__ __noinline__ uint32_t
iqr_threshold_exact(uint32_t sig_packed, uint32_t lane, int multiplier, uint32_t threshold)
{
  uint32_t sig8[8];
#pragma unroll
  for (int j = 0; j < 8; ++j)
    sig8[j] = __shfl_sync(0xFFFFFFFFu, sig_packed, int((lane & ~7u) + j));
  const uint64_t K = uint64_t(int64_t(int16_t(multiplier))) * uint64_t(threshold); // (sigma*mult)*thr mod 2^64
  int th[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int ch16 = int(2 * (lane & 7u)) + h;                           // channel within the 16-channel register
    const int pos = ch16 < 8 ? ch16 : (ch16 == 15 ? 8 : ch16 + 1);       // inverse of perm {0..7,15,8..14}
    const int g = pos >> 2, jj = pos & 3;
    uint64_t v = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int pp = 4 * g + i;
      const int cc = pp < 8 ? pp : (pp == 8 ? 15 : pp - 1);              // perm
      uint32_t w = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        w = (cc >> 1) == q ? sig8[q] : w;
      const uint64_t s16 = (cc & 1) ? (w >> 16) : (w & 0xFFFFu);
      v |= s16 << (16 * i);
    }
    v *= K;
    th[h] = int(int16_t(uint16_t(v >> (16 * jj))));
  }
  return pack2(th[0], th[1]);
}

struct ChanRegs
{ // one channel, unpacked
  int median, accum, prev, charge, tover, peak_adc, peak_time;
  int q25, q75, a25, a75;
  int rs, med_rs, acc_rs, rs_factor;
  int ring[8];
};

template<int ALGO, bool WIB2>
struct ScalarAlgo
{
  static constexpr int kGroupUnroll = 1;
  static constexpr int kWarpsPerSm = 0; // persistent single-warp CTAs per SM the WIBEth launch aims for; 0 = as many as fit
  static constexpr int kQuadCtasPerSm = 0; // WIBEth: > 0 = run wibeth_quad_kernel (CTA of 4 links + producer warp) with that many CTAs per SM
  static constexpr int kWib2MinCtas = 1;   // WIB2: minimum CTAs per SM the register allocation must allow
  ChanRegs c[2];
  uint32_t kphase; // FIR ring phase

  __device__ __forceinline__ void load(const uint32_t* st, uint32_t lane, uint32_t flags)
  {
    auto ld = [&](int v) { return st[v * 32 + lane]; };
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      auto g = [&](int v) { uint32_t w = ld(v); return h ? hi16s(w) : lo16s(w); };
      ChanRegs& r = c[h];
      r.median = g(SV_MEDIAN); r.accum = g(SV_ACCUM); r.prev = g(SV_PREV) & 0xFFFF; r.charge = g(SV_CHARGE) & 0xFFFF;
      r.tover = g(SV_TOVER) & 0xFFFF; r.peak_adc = g(SV_PEAK_ADC) & 0xFFFF; r.peak_time = g(SV_PEAK_TIME) & 0xFFFF;
      r.q25 = g(SV_Q25); r.q75 = g(SV_Q75); r.a25 = g(SV_A25); r.a75 = g(SV_A75);
      r.rs = g(SV_RS); r.med_rs = g(SV_MED_RS); r.acc_rs = g(SV_ACC_RS); r.rs_factor = g(SV_RS_FACTOR);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        r.ring[j] = g(SV_RING0 + j);
    }
    kphase = (flags >> 8) & 7u;
  }
  __device__ __forceinline__ void store(uint32_t* st, uint32_t lane, uint32_t) const
  {
    auto put = [&](int v, int lo, int hi) { st[v * 32 + lane] = pack2(lo, hi); };
    put(SV_MEDIAN, c[0].median, c[1].median); put(SV_ACCUM, c[0].accum, c[1].accum); put(SV_PREV, c[0].prev, c[1].prev);
    put(SV_CHARGE, c[0].charge, c[1].charge); put(SV_TOVER, c[0].tover, c[1].tover);
    put(SV_PEAK_ADC, c[0].peak_adc, c[1].peak_adc); put(SV_PEAK_TIME, c[0].peak_time, c[1].peak_time);
    put(SV_Q25, c[0].q25, c[1].q25); put(SV_Q75, c[0].q75, c[1].q75); put(SV_A25, c[0].a25, c[1].a25); put(SV_A75, c[0].a75, c[1].a75);
    put(SV_RS, c[0].rs, c[1].rs); put(SV_MED_RS, c[0].med_rs, c[1].med_rs); put(SV_ACC_RS, c[0].acc_rs, c[1].acc_rs);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      put(SV_RING0 + j, c[0].ring[j], c[1].ring[j]);
  }
  __device__ __forceinline__ uint32_t phase_after(uint32_t ticks) const { return (kphase + ticks) & 7u; }
  __device__ __forceinline__ void configure(const KernelParams&) {}
  static __device__ __forceinline__ uint32_t sample(const uint32_t* row, const PairPos& pp) { return extract_pair(row, pp); }
  template<int ROW_WORDS = 28>
  __device__ __forceinline__ void begin_chunk(const uint32_t*, const PairPos&) {}
  template<bool WIB2_UNITS = false>
  __device__ __forceinline__ void finish_link(const TickCtx&) {}
  template<bool WIB2_UNITS>
  static __device__ __forceinline__ void flush(const HitStage&, const TpSink&, const uint8_t*, uint32_t, uint32_t, bool = true) {} // emits directly

  // setState: pedestal = first sample, quartiles +-20 (wibeth/tpg/ProcessingInfo.hpp:116-144)
  __device__ __forceinline__ void seed(uint32_t S)
  {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int ped = h ? int(S >> 16) : int(S & 0xFFFFu);
      c[h].median = ped;
      c[h].q25 = wrap16(ped - 20);
      c[h].q75 = wrap16(ped + 20);
    }
  }

  template<int G, bool DUMP, int ROW_WORDS = 28, bool WIB2_UNITS = false>
  __device__ __forceinline__ void group(const uint32_t* rows, const PairPos& pp, const TickCtx& ctx, int t0, uint32_t* ped_out,
                                        uint32_t* wav_out, bool /*more*/)
  {
#pragma unroll
    for (int g = 0; g < G; ++g) {
      uint32_t ped, wav;
      tick(extract_pair(rows + g * ROW_WORDS, pp), ctx, t0 + g, (ctx.chan0 >> 1) & 31u, ped, wav);
      if constexpr (DUMP) {
        ped_out[g] = ped;
        wav_out[g] = wav;
      }
    }
  }

  // One tick for both channels of the lane. Returns {pedestal u16x2, waveform s16x2} through refs (debug dumps).
  __device__ __forceinline__ void tick(uint32_t S, const TickCtx& ctx, int t, uint32_t lane, uint32_t& ped_out, uint32_t& wav_out)
  {
    const KernelParams& p = *ctx.p;
    const int thr = int(int16_t(uint16_t(p.threshold)));
    int ped[2], wav[2];
    if constexpr (ALGO == SWTPG_ALGO_FIR_IQR) {
      // wib2/tpg/ProcessAVX2FIR.hpp:103-283
      const int multiplier = 1 << p.tap_exponent;
      const int adc_max = 32767 / multiplier;
      const int sigma_max = (1 << 15) / (multiplier * 5);
      int sigma[2], filt[2];
      const uint32_t kk = (kphase + ctx.tick_base + uint32_t(t)) & 7u; // absTimeModNTAPS at this tick
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        ChanRegs& r = c[h];
        const int raw = h ? int(S >> 16) : int(S & 0xFFFFu);
        const bool is_gt = raw > r.median, is_lt = raw < r.median;
        frugal_scalar(r.q25, raw, r.a25, 10, is_lt);
        frugal_scalar(r.q75, raw, r.a75, 10, is_gt);
        frugal_scalar(r.median, raw, r.accum, 10, true);
        int x = wrap16(raw - r.median);
        int sg = wrap16(r.q75 - r.q25);
        sg = sg > sigma_max ? sigma_max : sg;
        sigma[h] = sg;
        x = x > adc_max ? adc_max : x;
        int f = 0;
#pragma unroll
        for (int j = 0; j < 7; ++j) {
          int rv = 0; // ring[(j + kphase + t) & 7] without dynamic register indexing
#pragma unroll
          for (int q = 0; q < 8; ++q)
            rv = (((j + kk) & 7) == q) ? r.ring[q] : rv;
          f = wrap16(f + wrap16(p.taps[j] * rv));
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (kk == q)
            r.ring[q] = x;
        filt[h] = f;
        ped[h] = r.median;
        wav[h] = f;
      }
      // Threshold = 16-bit lane of a 64-bit-lane product over 4 adjacent AVX2 register POSITIONS.
      const uint32_t th2 = iqr_threshold_exact(pack2(sigma[0], sigma[1]), lane, multiplier, p.threshold);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int th = h ? hi16s(th2) : lo16s(th2);
        ChanRegs& r = c[h];
        const bool over = filt[h] > th;
        const bool left = r.prev && !over;
        r.charge = sat16(int(int16_t(r.charge)) + ((over ? filt[h] : 0) >> p.tap_exponent)) & 0xFFFF;
        r.tover = sat16(int(int16_t(r.tover)) + (over ? 1 : 0)) & 0xFFFF;
        if (left) {
          emit_wib2(p.sink, ctx.ts, t, uint32_t(r.charge), uint32_t(r.tover), ctx.chan0 + h, ctx.link);
          r.charge = r.tover = 0;
        }
        r.prev = over ? 0xFFFF : 0;
      }
    } else if constexpr (WIB2 && ALGO == SWTPG_ALGO_ABS_RS) {
      // wib2/tpg/ProcessRSAVX2.hpp:24-330: quartile trackers as in the FIR finder, running sum with the literal factors
      // R = 8 and scale = 5 (:29-33), every frugal limit 10, threshold sigma * info.threshold on 4 x int64 lanes (:198),
      // charge accumulates adds(RS, medianRS) >> tap_exponent (:210-213).
      const int multiplier = 1 << p.tap_exponent;
      const int sigma_max = (1 << 15) / (multiplier * int(p.threshold)); // threshold >= 1 (checked by the host)
      int sigma[2], lv[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        ChanRegs& r = c[h];
        const int raw = h ? int(S >> 16) : int(S & 0xFFFFu);
        const bool is_gt = raw > r.median, is_lt = raw < r.median;
        frugal_scalar(r.q25, raw, r.a25, 10, is_lt);
        frugal_scalar(r.q75, raw, r.a75, 10, is_gt);
        frugal_scalar(r.median, raw, r.accum, 10, true);
        const int x = wrap16(raw - r.median);
        const int first = wrap16(r.rs * 8);
        const int second = wrap16((x < 0 ? wrap16(-x) : x) * 5);
        int rs = wrap16((((wrap16(first + second) * 3276) >> 14) + 1) >> 1); // _mm256_mulhrs_epi16(sum, 32768/10)
        frugal_scalar(r.med_rs, rs, r.acc_rs, 10, true);
        rs = wrap16(rs - r.med_rs);
        r.rs = rs;
        lv[h] = rs;
        int sg = wrap16(r.q75 - r.q25);
        sigma[h] = sg > sigma_max ? sigma_max : sg;
        ped[h] = r.median;
        wav[h] = rs;
      }
      const uint32_t th2 = iqr_threshold_exact(pack2(sigma[0], sigma[1]), lane, 1, p.threshold);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        ChanRegs& r = c[h];
        const int th = h ? hi16s(th2) : lo16s(th2);
        const bool over = lv[h] > th;
        const bool left = r.prev && !over;
        const int temp = sat16(lv[h] + r.med_rs);
        r.charge = sat16(int(int16_t(r.charge)) + ((over ? temp : 0) >> p.tap_exponent)) & 0xFFFF;
        r.tover = sat16(int(int16_t(r.tover)) + (over ? 1 : 0)) & 0xFFFF;
        if (left) {
          emit_wib2(p.sink, ctx.ts, t, uint32_t(r.charge), uint32_t(r.tover), ctx.chan0 + h, ctx.link);
          r.charge = r.tover = 0;
        }
        r.prev = over ? 0xFFFF : 0;
      }
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        ChanRegs& r = c[h];
        const int raw = h ? int(S >> 16) : int(S & 0xFFFFu);
        const int L = (WIB2 && ALGO == SWTPG_ALGO_SIMPLE_THRESHOLD) ? 10 : p.acc_limit; // wib2/tpg/ProcessAVX2.hpp:79
        frugal_scalar(r.median, raw, r.accum, L, true);
        const int x = wrap16(raw - r.median);
        int level = x; // what the threshold sees
        if constexpr (ALGO == SWTPG_ALGO_ABS_RS || ALGO == SWTPG_ALGO_STANDARD_RS) {
          // wibeth/tpg/ProcessAbsRSAVX2.hpp:137-159, ProcessStandardRSAVX2.hpp:140-144
          const int first = wrap16(r.rs * int(int16_t(r.rs_factor)));
          const int ax = x < 0 ? wrap16(-x) : x;
          const int sum = ALGO == SWTPG_ALGO_STANDARD_RS ? wrap16(first + x) : wrap16(first + wrap16(ax * p.rs_scale));
          int rs = wrap16((((sum * 3276) >> 14) + 1) >> 1); // _mm256_mulhrs_epi16(sum, 32768/10)
          frugal_scalar(r.med_rs, rs, r.acc_rs, p.acc_limit, true);
          rs = wrap16(rs - r.med_rs);
          r.rs = rs;
          level = rs;
        }
        const bool over = level > thr;
        const bool left = r.prev && !over;
        if constexpr (WIB2) { // wib2/tpg/ProcessAVX2.hpp:104-121
          r.charge = sat16(int(int16_t(r.charge)) + ((over ? x : 0) >> p.tap_exponent)) & 0xFFFF;
          r.tover = sat16(int(int16_t(r.tover)) + (over ? 1 : 0)) & 0xFFFF;
          if (left) {
            emit_wib2(p.sink, ctx.ts, t, uint32_t(r.charge), uint32_t(r.tover), ctx.chan0 + h, ctx.link);
            r.charge = r.tover = 0;
          }
        } else { // wibeth/tpg/ProcessAVX2.hpp:114-204
          if constexpr (ALGO == SWTPG_ALGO_SIMPLE_THRESHOLD)
            r.charge = wrap16(int(int16_t(r.charge)) + (over ? x : 0)) & 0xFFFF; // add_epi16 wraps (H3)
          else
            r.charge = sat16(int(int16_t(r.charge)) + (over ? x : 0)) & 0xFFFF;  // adds_epi16
          if (x > int(int16_t(r.peak_adc))) { // not gated by `over` (H6)
            r.peak_adc = x & 0xFFFF;
            r.peak_time = r.tover;
          }
          r.tover = sat16(int(int16_t(r.tover)) + (over ? 1 : 0)) & 0xFFFF;
          if (left) {
            emit_wibeth(p.sink, ctx.ts, t, uint32_t(r.charge), uint32_t(r.tover), uint32_t(r.peak_adc), uint32_t(r.peak_time),
                        ctx.chan0 + h, ctx.link);
            r.charge = r.tover = r.peak_adc = r.peak_time = 0;
          }
        }
        r.prev = over ? 0xFFFF : 0;
        ped[h] = r.median;
        wav[h] = level;
      }
    }
    ped_out = pack2(ped[0], ped[1]);
    wav_out = pack2(wav[0], wav[1]);
  }
};

// =====================================================================================================================
// Packed fast path: SimpleThreshold on either frame layout (wibeth/tpg/ProcessAVX2.hpp:23-229, wib2/tpg/ProcessAVX2.hpp:24-200),
// two channels per 32-bit register.
// Validity (checked by the host before selecting it): 1 <= L <= 1000 and 0 <= threshold <= 32767.
//   * median m stays in [0, 16383] (it only ever steps towards samples, which are 14-bit), so adds_epi16 on it never
//     saturates;
//   * with a constant L >= 1 the accumulator is in [-L, L] between ticks, so "acc > L" <=> acc == L+1 and
//     "-acc > L" <=> acc == -(L+1);
//   * pedestal-subtracted samples s' are in [-16383, 16383] and threshold / peak_adc are non-negative, which is the
//     operand range gt2_mask_nonneg requires.
// Register forms:
//   Sb  = S - 16384           the sample as extract_pair_biased delivers it (bits 14, 15 forced to one), and
//   Mq  = 16385 - median      so that  Sb + Mq = s' + 1  ("sp1": the pedestal-subtracted sample, biased by one),
//                             clamp(Sb + Mq, 0, 2) = sign(s - m) + 1 is a single VIADDMNMX.S16x2.RELU, and Mq stays in
//                             [2, 16385]: it never passes through 0x0000 / 0xFFFF, so its +-1 steps are ONE 32-bit
//                             three-input add (IADD3) — no carry or borrow can cross the halves;
//   A   = accumulator - 1     as the bit pattern of an fp16x2 subnormal (value * 2^-24, sign-magnitude: exact integer
//                             arithmetic for |v| <= 1023 on the FMA pipe);
//   PK1 = peak_adc + 1        compared against sp1;   thr1 = threshold + 1 likewise;   Tn = -tover, PTn = -peak_time.
// The recurrence median(t) -> median(t+1) is the critical path of a link (time is sequential per channel): VIADDMNMX ->
// HADD2 -> HFMA2.SAT -> IADD3, four dependent instructions per tick (round 1: five). Everything else is taken off that
// path by a two-stage software pipeline inside group(): the raw words of the NEXT group's rows are loaded before this
// group's arithmetic starts (prefetch across the quiet-test branch), and the quiet test / hit bookkeeping of a group is
// DEFERRED by one group, so that it is scheduled in the shadow of the next group's pedestal chain.
// =====================================================================================================================
// PIPE = true: the two-stage software pipeline, the form every WIBEth SimpleThreshold launch runs (a warp runs at the speed of
// its dependent chain; since the accumulator reset is one IMAD that chain is short enough for 20 such warps per SM to beat 28
// straight-line ones); PIPE = false: straight-line groups, fewer registers and instructions — the WIB2 kernel's form and the
// base of the running-sum policies. The host picks per launch (launch_wibeth_simple, swtpg_capi.cu).
template<bool PIPE>
struct PackedSimpleT
{
  static constexpr int kGroupUnroll = SWTPG_GROUP_UNROLL;
  // persistent warps per SM, measured: 5 per sub-partition for the straight-line form (profiles/r01_warps_sweep.txt), 4 for the
  // pipelined one below one round of links (92 registers; profiles/r02_simple_pipeline_sweep.txt); from 2960 links on the host
  // asks for 20 (profiles/r02_pipelined_full_load_probe.txt, r02_warps_crossover_probe.txt)
  static constexpr int kWarpsPerSm = PIPE ? 16 : 20;
  static constexpr int kQuadCtasPerSm = SWTPG_QUAD_SIMPLE; // 0: one warp per CTA (measured faster for this policy)
  static constexpr int kWib2MinCtas = 5;
  static constexpr bool kWib2Fields = false; // which process_swtpg_hits derives the TP fields (and whether the peak is tracked)
  uint32_t Mq, A, prev, C, Tn, PK1, PTn;
  uint32_t cUp, cDn, thr1;
  uint32_t cL, cLn, c2L, cL1, cLL, kM1; // integer accumulator form: L, -L, (2L, 2L), L + 1, (L, L); 0xFFFFFFFF in a register
  uint32_t shift, shmask; // WIB2 flavour only
  uint32_t raw[8];        // software pipeline: words of the next group's four rows ...
  uint32_t dsp[4], dwhen; // ... and the deferred group: its four s' + 1 and (unit << 6 | first tick); dvalid below
  bool dvalid;

  static __device__ __forceinline__ uint32_t sample(const uint32_t* row, const PairPos& pp) { return extract_pair_biased(row, pp); }

  __device__ __forceinline__ void configure(const KernelParams& p)
  {
    const uint32_t L = uint32_t(p.acc_limit) & 0xFFFFu;
    const uint32_t up = L + 1u;            // acc == L + 1       (as an fp16 subnormal: value * 2^-24 == bit pattern)
    const uint32_t dn = 0x8000u | L;       // -L: addend of the "acc - L" / "-acc - L" saturating tests (sign-magnitude)
    cUp = up | (up << 16);
    cDn = dn | (dn << 16);
    cL = L;
    cLn = p.acc_limit_neg;
    cL1 = L + 1u;
    cLL = L * 0x00010001u;
    c2L = 2u * L * 0x00010001u;
    kM1 = p.all_ones;
    uint32_t th = p.threshold > 16383u ? 16383u : p.threshold; // s' <= 16383: any larger threshold is never exceeded
    th += 1u;
    thr1 = th | (th << 16);
    shift = shmask = 0;
  }
  // Register form of the accumulator: (acc - 1) as the bit pattern of an fp16x2 subnormal (value * 2^-24, sign-magnitude).
  static __device__ __forceinline__ uint32_t acc_to_reg(uint32_t v)
  {
    v = add2(v, 0xFFFFFFFFu);
    const uint32_t neg = (v & 0x80008000u) >> 15; // 1 per negative half
    const uint32_t m = neg * 0xFFFFu;             // 0xFFFF per negative half
    return (add2(v ^ m, neg) & 0x7FFF7FFFu) | (m & 0x80008000u);
  }
  static __device__ __forceinline__ uint32_t acc_from_reg(uint32_t v)
  {
    const uint32_t neg = (v & 0x80008000u) >> 15;
    const uint32_t m = neg * 0xFFFFu;
    return add2(add2((v & 0x7FFF7FFFu) ^ m, neg), 0x00010001u);
  }
  __device__ __forceinline__ void load(const uint32_t* st, uint32_t lane, uint32_t)
  {
    Mq = add2(~st[SV_MEDIAN * 32 + lane], 0x40024002u); // ~m = -m - 1
#if SWTPG_SIMPLE_INT
    A = add2(st[SV_ACCUM * 32 + lane], cLL);            // acc + L in [0, 2L]
#else
    A = acc_to_reg(st[SV_ACCUM * 32 + lane]);
#endif
    prev = st[SV_PREV * 32 + lane];
    C = st[SV_CHARGE * 32 + lane];
    Tn = neg2(st[SV_TOVER * 32 + lane]);
    PK1 = add2(st[SV_PEAK_ADC * 32 + lane], 0x00010001u);
    PTn = neg2(st[SV_PEAK_TIME * 32 + lane]);
    dvalid = false;
    dwhen = 0u;
#pragma unroll
    for (int g = 0; g < 4; ++g)
      dsp[g] = 0u; // s' = -1: below every threshold and every PK1, so an empty pipeline stage changes nothing
  }
  __device__ __forceinline__ uint32_t median() const { return add2(~Mq, 0x40024002u); } // 16385 - Mq
  __device__ __forceinline__ void store(uint32_t* st, uint32_t lane, uint32_t) const
  {
    st[SV_MEDIAN * 32 + lane] = median();
#if SWTPG_SIMPLE_INT
    st[SV_ACCUM * 32 + lane] = add2(A, neg2(cLL));
#else
    st[SV_ACCUM * 32 + lane] = acc_from_reg(A);
#endif
    st[SV_PREV * 32 + lane] = prev;
    st[SV_CHARGE * 32 + lane] = C;
    st[SV_TOVER * 32 + lane] = neg2(Tn);
    st[SV_PEAK_ADC * 32 + lane] = add2(PK1, 0xFFFFFFFFu);
    st[SV_PEAK_TIME * 32 + lane] = neg2(PTn);
  }
  __device__ __forceinline__ uint32_t phase_after(uint32_t) const { return 0; }
  __device__ __forceinline__ void seed(uint32_t Sb) { Mq = add2(~Sb, 0x00020002u); } // 16385 - S = 1 - Sb
  template<bool WIB2_UNITS>
  static __device__ __forceinline__ void flush(const HitStage& h, const TpSink& k, const uint8_t* link_base, uint32_t link, uint32_t lane,
                                               bool everything = true)
  {
    h.template flush<WIB2_UNITS, false>(k, link_base, link, lane, everything);
  }

  // frugal streaming median (wibeth/tpg/UtilsAVX2.hpp:38-73) + pedestal subtraction (ProcessAVX2.hpp:85): Sb -> s' + 1.
  // The accumulator lives as an fp16x2 subnormal stored minus one, so that adding sign+1 lands on the new value; both step
  // flags come out of the FMA pipe as the integer bit patterns 0 / 1 (saturating fp16 FMAs of +-acc - L), and the median
  // takes them in one 32-bit add. The ALU pipe does the sign, one compare and that add.
  //
  // Integer form (SWTPG_SIMPLE_INT, round 2): the accumulator biased by +L is a small non-negative number per half, like the
  // sign code, so their sum is a plain 32-bit add (either pipe); one clamp takes out the two step values, the difference is
  // the step D in {-1, 0, +1} per half held as ONE 32-bit integer (the low half may borrow from the high one — harmless, only
  // linear 32-bit operations consume it and their results have non-negative halves again), and the accumulator returns to L
  // where it stepped through one IMAD. Two instructions on the ALU pipe, two (one without the final add) on the FMA pipe, three
  // that issue on either; the fp16 form has 2 / 5 / 1.
  __device__ __forceinline__ uint32_t pedestal_step(uint32_t Sb)
  {
    const uint32_t sg1 = addclamp2(Sb, Mq, 0x00020002u);        // sign(s - m) + 1 in {0,1,2}
#if SWTPG_SIMPLE_INT
    const uint32_t U = A + sg1;                                 // acc' + L + 1 in [0, 2L + 2]
    const uint32_t c = addclamp2(U, kM1, c2L);                  // clamp(acc' + L, 0, 2L)
#if SWTPG_SIMPLE_INT == 2
    const uint32_t W = Mq - U + 0x00010001u;                    // in the shadow of the clamp
    const uint32_t V = cLL - cL * U;                            // L - L U
    Mq = W + c;                                                 // m += U - 1 - c
    A = cL1 * c + V;                                            // c - L (U - 1 - c)
#else
    const uint32_t D = U - c - 0x00010001u;                     // the step
#if SWTPG_STEP_NEGL
    A = cLn * D + c;                                            // back to L where it stepped: c - L D as ONE IMAD with -L from the
                                                                // constant bank (written c - L * D ptxas negates D first: 62
                                                                // instead of 59 instructions per quiet 4-tick group)
#else
    A = c - cL * D;
#endif
    Mq -= D;                                                    // (steering these adds to the FMA pipe measured 1 % slower here)
#endif
    return add2(Sb, Mq);                                        // s' + 1 with the UPDATED median
#else
    const uint32_t T = hadd2_bits(A, sg1);                      // acc after this sample, in [-(L+1), L+1]
    const uint32_t up1 = hfma2_sat_bits(T, 0x3C003C00u, cDn);   // sat(acc - L)  -> bit pattern 1: acc == L+1
    const uint32_t dn1 = hfma2_sat_bits(T, 0xBC00BC00u, cDn);   // sat(-acc - L) -> bit pattern 1: acc == -(L+1)
    const uint32_t keep = ne2_abs_one(T, cUp);                  // 1.0 unless |acc| == L+1
    A = hfma2_bits(keep, T, 0x80018001u);                       // (stepped ? 0 : acc) - 1
    Mq = Mq + dn1 - up1;                                        // m += up - down: one IADD3, halves cannot interact (see above)
    return add2(Sb, Mq);                                        // s' + 1 with the UPDATED median
#endif
  }
  __device__ __forceinline__ uint32_t over_mask(uint32_t sp1) const { return gt2_mask_nonneg(sp1, thr1); } // (:97-98)

  // Hit bookkeeping of one tick (:102-207). A lane on which a channel ends a hit parks its packed registers in the warp's
  // staging buffer and resets (lane-divergent); accepted iff hit_charge != 0 (src/wibeth/WIBEthFrameProcessor.cpp:520),
  // which flush() decides on the masked charge.
  __device__ __forceinline__ void hit_update(uint32_t sp1, const TickCtx& ctx, uint32_t unit, uint32_t t)
  {
    const uint32_t over = over_mask(sp1);
    const uint32_t left = prev & ~over;                 //                                  (:102)
    C = add2(C, add2(sp1, 0xFFFFFFFFu) & over);         // wrapping charge                  (:114-118)
    const uint32_t gtp = gt2_mask_nonneg(sp1, PK1);     // un-gated peak tracking           (:134-136)
    PK1 = max2(PK1, sp1);
    PTn = (Tn & gtp) | (PTn & ~gtp);                    // peak_time = tover BEFORE increment
    Tn = addmax2(Tn, over, 0x80018001u);                // tover = adds(tover, 1): -tover >= -32767   (:139-140)
    prev = over;
    if (left != 0u) {                                   //                                  (:154-204)
      ctx.stage->push(HitStage::meta(ctx.chan0, unit, t), C & left, neg2(Tn), add2(PK1, 0xFFFFFFFFu), neg2(PTn));
      C &= ~left;
      Tn &= ~left;
      PK1 = (PK1 & ~left) | (left & 0x00010001u);
      PTn &= ~left;
    }
  }

  // Quiet test and hit bookkeeping of four ticks, two tiers (results identical in both):
  //  QUIET — no channel of the warp is inside a hit and none goes over threshold in the group (by far the most common case on
  //     physical noise): per channel charge = tover = peak_time = 0 stay 0 and only the un-gated peak tracker moves,
  //     peak_adc = max(peak_adc, max_g s'_g), because with tover == 0 every peak update writes peak_time = 0 again
  //     (ProcessAVX2.hpp:134-136). Outside a hit peak_adc <= threshold (it restarts from 0 when a hit ends and every sample
  //     since was not over), so "new peak > threshold" <=> "some sample of the group is over".
  //  otherwise per-tick bookkeeping for the whole warp.
  template<bool WIB2_UNITS>
  __device__ __forceinline__ void hits_of_group(const uint32_t (&sp)[4], const TickCtx& ctx, uint32_t unit, uint32_t t0, bool valid)
  {
    const uint32_t pk = __vimax3_s16x2(__vimax3_s16x2(sp[0], sp[1], sp[2]), sp[3], PK1);
    const uint32_t busy = over_mask(pk) | prev; // some tick of the group over threshold, or still inside a hit
    if (__builtin_expect(!__any_sync(0xFFFFFFFFu, busy != 0u), 1)) {
      PK1 = pk;
      return;
    }
    if (!valid) // empty pipeline stage (first group of a link) while a hit carried over from the previous batch is still open
      return;
#pragma unroll
    for (int g = 0; g < 4; ++g)
      hit_update(sp[g], ctx, unit, t0 + uint32_t(g));
    if (ctx.stage->must_flush())
      flush<WIB2_UNITS>(*ctx.stage, ctx.p->sink, ctx.link_base, ctx.link, (ctx.chan0 >> 1) & 31u, false); // whole 32-pair rounds only
  }

  // First group of a chunk: its rows' words come straight from shared memory (the copy has just landed).
  template<int ROW_WORDS = 28>
  __device__ __forceinline__ void begin_chunk(const uint32_t* rows, const PairPos& pp)
  {
    if constexpr (PIPE) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        raw[2 * g] = rows[g * ROW_WORDS + pp.w0];
        raw[2 * g + 1] = rows[g * ROW_WORDS + pp.w1];
      }
    }
  }

  // G consecutive ticks. `more`: another group of the same chunk follows (its rows are prefetched before the branch).
  template<int G, bool DUMP, int ROW_WORDS = 28, bool WIB2_UNITS = false>
  __device__ __forceinline__ void group(const uint32_t* rows, const PairPos& pp, const TickCtx& ctx, int t0, uint32_t* ped_out,
                                        uint32_t* wav_out, bool more)
  {
    static_assert(G == 4, "max tree and pipeline registers are written for 4 ticks");
    uint32_t sp[G], Sb[G];
    if constexpr (PIPE) {
#pragma unroll
      for (int g = 0; g < G; ++g)
        Sb[g] = extract_pair_biased_from(raw[2 * g], raw[2 * g + 1], pp.sh);
      if (more) // the loads of the next group are in flight while this group's recurrence runs
        begin_chunk<ROW_WORDS>(rows + G * ROW_WORDS, pp);
    } else {
#pragma unroll
      for (int g = 0; g < G; ++g)
        Sb[g] = extract_pair_biased(rows + g * ROW_WORDS, pp);
    }
    // The pedestal recurrence never reads hit state: one branch-free dependent chain.
#pragma unroll
    for (int g = 0; g < G; ++g) {
      sp[g] = pedestal_step(Sb[g]);
      if constexpr (DUMP) {
        ped_out[g] = median();
        wav_out[g] = add2(sp[g], 0xFFFFFFFFu);
      }
    }
    if constexpr (PIPE) {
      // ... and in its shadow the quiet test (max tree, compare, vote) of the PREVIOUS group, which is independent of it
      hits_of_group<WIB2_UNITS>(dsp, ctx, dwhen >> 6, dwhen & 63u, dvalid);
#pragma unroll
      for (int g = 0; g < G; ++g)
        dsp[g] = sp[g];
      dwhen = (ctx.unit << 6) | uint32_t(t0);
      dvalid = true;
    } else {
      hits_of_group<WIB2_UNITS>(sp, ctx, ctx.unit, uint32_t(t0), true);
    }
  }
  // End of a link (before the staged hits are flushed and the state is stored): drain the pipeline.
  template<bool WIB2_UNITS = false>
  __device__ __forceinline__ void finish_link(const TickCtx& ctx)
  {
    if constexpr (PIPE) {
      if (dvalid)
        hits_of_group<WIB2_UNITS>(dsp, ctx, dwhen >> 6, dwhen & 63u, true);
      dvalid = false;
    }
  }
};
using PackedSimpleWibEth = PackedSimpleT<false>;
using PackedSimpleWibEthPipe = PackedSimpleT<true>;

// =====================================================================================================================
// WIB2 SimpleThreshold (wib2/tpg/ProcessAVX2.hpp:24-200). Same pedestal recurrence with the limit fixed at 10 (:79); charge
// accumulates (over ? s' : 0) >> tap_exponent with signed saturation (:110-112), no peak tracking, hit block = {chan, t,
// charge, tover}. Validity: 0 <= threshold <= 32767. Straight-line groups only: the CTA form caps the registers for five CTAs
// per SM, and the pipelined form measured 7 % slower there (gpurun_out/r02_probe3.txt).
// =====================================================================================================================
struct PackedSimpleWib2 : PackedSimpleWibEth
{
  static constexpr bool kWib2Fields = true;
  template<bool WIB2_UNITS>
  static __device__ __forceinline__ void flush(const HitStage& h, const TpSink& k, const uint8_t* link_base, uint32_t link, uint32_t lane,
                                               bool everything = true)
  {
    h.template flush<WIB2_UNITS, true>(k, link_base, link, lane, everything);
  }

  __device__ __forceinline__ void configure(const KernelParams& p)
  {
    KernelParams q = p;
    q.acc_limit = 10; // wib2/tpg/ProcessAVX2.hpp:79
    q.acc_limit_neg = 0u - 10u;
    PackedSimpleWibEth::configure(q);
    shift = uint32_t(p.tap_exponent);
    const uint32_t m = 0xFFFFu >> shift;
    shmask = m | (m << 16);
  }
  __device__ __forceinline__ void hit_update(uint32_t sp1, const TickCtx& ctx, uint32_t unit, uint32_t t)
  {
    const uint32_t over = over_mask(sp1);
    const uint32_t left = prev & ~over;
    // over => s' > threshold >= 0: both halves of the masked value are non-negative, so the arithmetic shift is a logical one
    const uint32_t add = ((add2(sp1, 0xFFFFFFFFu) & over) >> shift) & shmask;
    C = minu2(add2(C, add), 0x7FFF7FFFu);               // adds_epi16 of two non-negative halves
    Tn = addmax2(Tn, over, 0x80018001u);                // tover = adds(tover, 1)
    prev = over;
    if (left != 0u) {
      // accepted iff hit_charge != 0 (src/wib2/WIB2FrameProcessor.cpp:429): decided in flush on the masked charge
      ctx.stage->push(HitStage::meta(ctx.chan0, unit, t), C & left, neg2(Tn));
      C &= ~left;
      Tn &= ~left;
    }
  }
  template<bool WIB2_UNITS>
  __device__ __forceinline__ void hits_of_group(const uint32_t (&sp)[4], const TickCtx& ctx, uint32_t unit, uint32_t t0, bool valid)
  {
    const uint32_t mx = __vimax3_s16x2(__vimax3_s16x2(sp[0], sp[1], sp[2]), sp[3], sp[3]);
    const uint32_t busy = over_mask(mx) | prev;
    if (__builtin_expect(!__any_sync(0xFFFFFFFFu, busy != 0u), 1))
      return; // nothing but the pedestal moves outside hits
    if (!valid)
      return;
#pragma unroll
    for (int g = 0; g < 4; ++g)
      hit_update(sp[g], ctx, unit, t0 + uint32_t(g));
    if (ctx.stage->must_flush())
      flush<WIB2_UNITS>(*ctx.stage, ctx.p->sink, ctx.link_base, ctx.link, (ctx.chan0 >> 1) & 31u, false); // whole 32-pair rounds only
  }
  template<int G, bool DUMP, int ROW_WORDS, bool WIB2_UNITS>
  __device__ __forceinline__ void group(const uint32_t* rows, const PairPos& pp, const TickCtx& ctx, int t0, uint32_t* ped_out,
                                        uint32_t* wav_out, bool /*more*/)
  {
    static_assert(G == 4, "max tree is written for 4 ticks");
    uint32_t sp[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      sp[g] = pedestal_step(extract_pair_biased(rows + g * ROW_WORDS, pp));
      if constexpr (DUMP) {
        ped_out[g] = median();
        wav_out[g] = add2(sp[g], 0xFFFFFFFFu);
      }
    }
    hits_of_group<WIB2_UNITS>(sp, ctx, ctx.unit, uint32_t(t0), true);
  }
};

// =====================================================================================================================
// Packed fast path: WIBEth running-sum finders — AbsRS (wibeth/tpg/ProcessAbsRSAVX2.hpp:21-345) and StandardRS
// (wibeth/tpg/ProcessStandardRSAVX2.hpp:23-). Per tick and channel:
//   s' = raw - median                                     (frugal pedestal as in SimpleThreshold)
//   sum = RS * R + |s'| * scale      (AbsRS, :137-150)    or   RS * R + s'   (StandardRS, :140-144)   — mullo / add, mod 2^16
//   RS  = mulhrs(sum, 32768/10)                           (UtilsAVX2.hpp:77-81: round(sum / 10))
//   RS -= median_RS  after a second frugal update on RS   (:152-159)
//   threshold on RS; charge accumulates s' with signed saturation (:204); peak tracking on s' as in SimpleThreshold.
// The two 16-bit multiplies are done per half with 32-bit IMADs (only the low 16 bits of a product are used, so the packed
// register itself is the low-half operand and no sign extension is needed); mulhrs is sign-extend + IMAD + shift per half.
// |RS| <= 3277 by construction, so the RS median lives in [-3277, 3277] and the packed sign tests cannot overflow.
// Validity: 1 <= L <= 1000, 0 <= threshold <= 32767; any per-channel memory factor and any scale factor.
// =====================================================================================================================
template<bool STANDARD>
struct PackedRsWibEth : PackedSimpleWibEth
{
  static constexpr int kGroupUnroll = SWTPG_RS_GROUP_UNROLL;
  static constexpr int kWarpsPerSm = STANDARD ? 20 : 16; // AbsRS: 4 warps per sub-partition beat 5 by 3 % (same sweep)
  uint32_t RS1, MRq, AR;     // RS + 1 (carried value, after median subtraction); 1 - median_RS; (acc_RS - 1) as fp16 subnormal
  int f_lo, f_hi, nf_lo, nf_hi, scale;
  template<int ROW_WORDS = 28>
  __device__ __forceinline__ void begin_chunk(const uint32_t*, const PairPos&) {}
  template<bool WIB2_UNITS = false>
  __device__ __forceinline__ void finish_link(const TickCtx&) {}

  __device__ __forceinline__ void configure(const KernelParams& p)
  {
    PackedSimpleWibEth::configure(p);
    // RS - median_RS is in [-6554, 6554]: a larger threshold is never exceeded (and stays inside the comparator's range)
    uint32_t th = p.threshold > 6555u ? 6555u : p.threshold;
    th += 1u;
    thr1 = th | (th << 16);
    scale = p.rs_scale;
  }
  __device__ __forceinline__ void load(const uint32_t* st, uint32_t lane, uint32_t flags)
  {
    PackedSimpleWibEth::load(st, lane, flags);
    RS1 = add2(st[SV_RS * 32 + lane], 0x00010001u);
    MRq = add2(~st[SV_MED_RS * 32 + lane], 0x00020002u);
    AR = acc_to_reg(st[SV_ACC_RS * 32 + lane]);
    const uint32_t f = st[SV_RS_FACTOR * 32 + lane];
    f_lo = lo16s(f);
    f_hi = hi16s(f);
    nf_lo = -f_lo; // RS * R = (RS + 1) * R - R: the "- R" rides on an IMAD addend
    nf_hi = -f_hi;
  }
  __device__ __forceinline__ void store(uint32_t* st, uint32_t lane, uint32_t k) const
  {
    PackedSimpleWibEth::store(st, lane, k);
    st[SV_RS * 32 + lane] = add2(RS1, 0xFFFFFFFFu);
    st[SV_MED_RS * 32 + lane] = add2(~MRq, 0x00020002u);
    st[SV_ACC_RS * 32 + lane] = acc_from_reg(AR);
  }
  // frugal update of (Mq_, A_) = (1 - median, accumulator - 1) with sample S; returns S - median + 1 with the updated median.
  // Used for the median of the RUNNING SUM, which lives in [-3277, 3277]: its register passes through zero, so the two step
  // flags are added as packed halves (two VIADD.16x2) instead of the one 32-bit add the raw pedestal gets away with.
  __device__ __forceinline__ uint32_t frugal(uint32_t S, uint32_t& Mq_, uint32_t& A_) const
  {
    const uint32_t sg1 = addclamp2(S, Mq_, 0x00020002u);
    const uint32_t T = hadd2_bits(A_, sg1);
    const uint32_t upm = eq2_mask(T, cUp);
    const uint32_t dn1 = hfma2_sat_bits(T, 0xBC00BC00u, cDn);
    A_ = hfma2_bits(ne2_abs_one(T, cUp), T, 0x80018001u);
    Mq_ = add2(add2(Mq_, upm), dn1);
    return add2(S, Mq_);
  }
  // _mm256_mulhrs_epi16(v, 3276) for the 16-bit value in the LOW half of `w` (upper half ignored): ((v * 3276 >> 14) + 1) >> 1
  // = (v * 3276 + 2^14) >> 15, which fits 32-bit arithmetic (|v * 3276| < 2^27).
  static __device__ __forceinline__ int mulhrs_low(uint32_t w)
  {
    return (int(int16_t(uint16_t(w))) * 3276 + 16384) >> 15;
  }
  // One tick: sp1 = s' + 1 (pedestal-subtracted sample, biased), returns RS - median_RS + 1
  __device__ __forceinline__ uint32_t rs_step(uint32_t sp1)
  {
    const uint32_t x = add2(sp1, 0xFFFFFFFFu);
    uint32_t lo, hi;
    if constexpr (STANDARD) {
      lo = uint32_t(int(RS1) * f_lo + int(x) + nf_lo);                       // low 16 bits: RS * R + s'
      hi = uint32_t(int(RS1 >> 16) * f_hi + int(x >> 16) + nf_hi);
    } else {
      const uint32_t ax = max2(x, neg2(x));                                  // |s'| (s' > -32768)
      lo = uint32_t(int(RS1) * f_lo + (int(ax) * scale + nf_lo));            // low 16 bits: RS * R + |s'| * scale
      hi = uint32_t(int(RS1 >> 16) * f_hi + (int(ax >> 16) * scale + nf_hi));
    }
    const int r_lo = mulhrs_low(lo), r_hi = mulhrs_low(hi);
    const uint32_t rs = __byte_perm(uint32_t(r_lo), uint32_t(r_hi), 0x5410);  // pack the two low halves
    RS1 = frugal(rs, MRq, AR);                                               // second pedestal, on the running sum
    return RS1;
  }

  __device__ __forceinline__ void hit_update(uint32_t sp1, uint32_t lv1, const TickCtx& ctx, int t)
  {
    const uint32_t over = gt2_mask_nonneg(lv1, thr1);
    const uint32_t left = prev & ~over;
    const uint32_t xm = add2(sp1, 0xFFFFFFFFu) & over;
    // adds_epi16 (:204); s' may be negative here. Packed wrapping add, then repair the (rare) halves that overflowed:
    // overflow <=> both operands have the same sign and the sum's sign differs.
    const uint32_t Cw = add2(C, xm);
    const uint32_t ovf = ~(C ^ xm) & (C ^ Cw) & 0x80008000u;
    if (__builtin_expect(ovf != 0u, 0))
      C = pack2(sat16(lo16s(C) + lo16s(xm)), sat16(hi16s(C) + hi16s(xm)));
    else
      C = Cw;
    const uint32_t gtp = gt2_mask_nonneg(sp1, PK1);                          // un-gated peak tracking on s'
    PK1 = max2(PK1, sp1);
    PTn = (Tn & gtp) | (PTn & ~gtp);
    Tn = addmax2(Tn, over, 0x80018001u);
    prev = over;
    if (left != 0u) {
      ctx.stage->push(HitStage::meta(ctx.chan0, ctx.unit, uint32_t(t)), C & left, neg2(Tn), add2(PK1, 0xFFFFFFFFu), neg2(PTn));
      C &= ~left;
      Tn &= ~left;
      PK1 = (PK1 & ~left) | (left & 0x00010001u);
      PTn &= ~left;
    }
  }

  template<int G, bool DUMP, int ROW_WORDS = 28, bool WIB2_UNITS = false>
  __device__ __forceinline__ void group(const uint32_t* rows, const PairPos& pp, const TickCtx& ctx, int t0, uint32_t* ped_out,
                                        uint32_t* wav_out, bool /*more*/)
  {
    static_assert(G == 4, "max trees below are written for 4 ticks");
    uint32_t sp[G], lv[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      sp[g] = pedestal_step(extract_pair_biased(rows + g * ROW_WORDS, pp));
      lv[g] = rs_step(sp[g]);
      if constexpr (DUMP) {
        ped_out[g] = median();
        wav_out[g] = add2(lv[g], 0xFFFFFFFFu);
      }
    }
    const uint32_t mxl = __vimax3_s16x2(__vimax3_s16x2(lv[0], lv[1], lv[2]), lv[3], lv[3]);
    const uint32_t busy = gt2_mask_nonneg(mxl, thr1) | prev;
    if (__builtin_expect(!__any_sync(0xFFFFFFFFu, busy != 0u), 1)) {
      // outside hits only the un-gated peak tracker moves (peak_time = tover = 0 is rewritten with 0)
      PK1 = __vimax3_s16x2(__vimax3_s16x2(sp[0], sp[1], sp[2]), sp[3], PK1);
      return;
    }
#pragma unroll
    for (int g = 0; g < G; ++g)
      hit_update(sp[g], lv[g], ctx, t0 + g);
    if (ctx.stage->must_flush())
      flush<WIB2_UNITS>(*ctx.stage, ctx.p->sink, ctx.link_base, ctx.link, (ctx.chan0 >> 1) & 31u, false); // whole 32-pair rounds only
  }
};

// =====================================================================================================================
// Packed fast path: FIR matched filter + IQR threshold (wib2/tpg/ProcessAVX2FIR.hpp:21-314), on either frame layout.
// Three frugal trackers per channel (quartiles on the lanes below / above the OLD median, then the median, all with
// L = 10, :108-125), sigma = min(q75 - q25, sigmaMax), s' = min(s - median, adcMax), the 7-tap filter over the ring
// (which excludes the two newest samples, :160-201), threshold sigma * multiplier * threshold, charge += filt >> exponent.
//
// Validity (checked by the host): taps == {1,6,15,20,15,6,1,0}, 1 <= tap_exponent <= 10, sigmaMax * multiplier * threshold
// < 2^16. What makes it fast:
//   * the taps firwin_int(7, 0.1, 64) produces are the binomial row (1+z)^6, so the filter is six cascaded two-tap adders:
//     6 packed adds per tick instead of 7 multiply-adds on a rotating ring, and every add wraps mod 2^16 exactly like the
//     reference's mullo/add chain (ring homomorphism). The carried state stays the reference's ring: it is converted to
//     cascade registers when a link is loaded and back when it is stored, which also removes the ring phase from the loop;
//   * accumulators are fp16x2 subnormals and the median lives as 16385 - median against the biased sample Sb = S - 16384,
//     with the four-instruction step of PackedSimpleT;
//   * the kernel is bound by the ALU pipe (LOP3 / SHF / VIADD.16x2 / VIADDMNMX / IADD3: 16 lanes per sub-partition, half the
//     issue rate), while the fp16 pipe (HFMA2 / HADD2 / HSET2) idles at 40 % — measured, profiles/r02_wibeth_fir_*. So every
//     tracker keeps its ALU work to the sign test and one three-input add: the step flags come out of the fp16 pipe as the
//     integer bit patterns 0 / 1 (saturating FMAs of +-acc - L), the quartile enables (sample below / above the OLD median,
//     :119-121) multiply into the accumulator's FMA, and q25 lives as 16385 - q25, q75 as q75 + 2 — both inside [2, 16405] for
//     every input, so their +-1 steps cannot carry between the halves. (Merging the two quartile trackers into one step on
//     per-half selected operands — they never move in the same tick — saves 6 instructions per tick but ADDS three to the
//     ALU pipe, the selects and write-backs being LOP3s: measured, no gain, not kept.)
//   * sigma is only needed where the threshold is: the quiet test works with a lower bound of the group's sigmas taken from the
//     quartiles after its last tick (sigma moves by at most one per tick, because at most one quartile moves), and the per-tick
//     values are formed from the saved quartile registers only in a group that turns out busy;
//   * while every sigma of the warp is >= 0 the 64-bit-lane product equals the per-channel product (no carries between
//     positions) and is one packed IMAD; a warp that sees a negative sigma in a 4-tick group recomputes that
//     group's thresholds with iqr_threshold_exact.
// =====================================================================================================================
struct PackedFirIqr
{
  static constexpr int kGroupUnroll = SWTPG_FIR_GROUP_UNROLL;
  static constexpr int kWarpsPerSm = 12; // 3 warps per sub-partition: 3 % faster than the 16 its 128 registers allow (same sweep)
  static constexpr int kWib2MinCtas = 5;
  static constexpr int kQuadCtasPerSm = 4; // CTA form, 16 consumer warps per SM (measured: 4 CTAs 29.8 %, 5 CTAs 27.8 %, profiles/r02_fir_forms.txt)
                                           // (64 registers instead of 109; profiles/r01_quad_vs_warp.txt)
#ifndef SWTPG_FIR_INT
#define SWTPG_FIR_INT 1 // 1: integer accumulators (round 2, see track()); 0: the fp16-subnormal forms of round 1 below
#endif
#ifndef SWTPG_FIR_U3
#define SWTPG_FIR_U3 0 // integer form: the two three-input adds as one IADD3 each (ALU pipe, 0) or as two IMADs each (FMA pipe, 1)
#endif
#ifndef SWTPG_FIR_MERGED
#define SWTPG_FIR_MERGED 1 // fp16 forms only. 1: the two quartile trackers share one step on per-half selected operands; 0: two separate steps
#endif
  // 16385 - median, 16385 - q25 in every form. Integer form: Q75p = 16385 - q75, A = acc + 10, A25 = acc25 + 10, A75 = acc75 + 11.
  // fp16 forms: Q75p = q75 + 2, A = acc - 1 and the quartile accumulators as fp16x2 subnormal bit patterns
  // (merged form acc25 - 1 and -acc75 - 1, separate form acc25 and acc75)
  uint32_t Mq, A, Q25n, A25, Q75p, A75;
  uint32_t kM1, kP1; // 0xFFFFFFFF and 1 in registers the compiler cannot see through (KernelParams::all_ones, one)
  uint32_t d1, d2, d3, d4, d5, d6, o1, o2; // cascade: d_j = previous input of stage j, o1/o2 = previous two outputs
  uint32_t prev, C, Tn;
  uint32_t xmax, sig3max, K, Kneg3, shift, shmask, thr_cfg, mult;

  static __device__ __forceinline__ uint32_t sample(const uint32_t* row, const PairPos& pp) { return extract_pair_biased(row, pp); }
  template<int ROW_WORDS = 28>
  __device__ __forceinline__ void begin_chunk(const uint32_t*, const PairPos&) {}
  template<bool WIB2_UNITS = false>
  __device__ __forceinline__ void finish_link(const TickCtx&) {}
  static constexpr uint32_t kTiny = 0x00010001u;    // 2^-24 per half
  static constexpr uint32_t kNegTiny = 0x80018001u;
  static constexpr uint32_t kUp = 0x000B000Bu;      // (L+1) * 2^-24, L = 10
  static constexpr uint32_t kDnEq = 0x800B800Bu;    // -(L+1)
  static constexpr uint32_t kNegL = 0x800A800Au;    // -L
  static constexpr uint32_t kOne = 0x3C003C00u, kNegOne = 0xBC00BC00u;

  template<bool WIB2_UNITS>
  static __device__ __forceinline__ void flush(const HitStage& h, const TpSink& k, const uint8_t* link_base, uint32_t link, uint32_t lane,
                                               bool everything = true)
  {
    h.template flush<WIB2_UNITS, true>(k, link_base, link, lane, everything); // FIR hit blocks carry {chan, t, charge, tover} only
  }
  __device__ __forceinline__ void configure(const KernelParams& p)
  {
    mult = 1u << p.tap_exponent;
    const uint32_t adc_max = 32767u / mult;                   // wib2/tpg/ProcessingInfo.hpp:93
    const uint32_t sigma_max = (1u << 15) / (mult * 5u);      // ProcessAVX2FIR.hpp:36
    xmax = adc_max * 0x00010001u;
    sig3max = (sigma_max + 3u) * 0x00010001u;
    thr_cfg = p.threshold;
    K = mult * p.threshold;
    Kneg3 = 0u - 3u * K * 0x00010001u;                        // -(3K, 3K) as ONE 32-bit integer (see threshold())
    shift = uint32_t(p.tap_exponent);
    const uint32_t m = 0xFFFFu >> shift;
    shmask = m | (m << 16);
    kM1 = p.all_ones;
    kP1 = p.one;
  }
  static __device__ __forceinline__ uint32_t to_sm(uint32_t v)
  { // two's complement s16x2 -> sign-magnitude (= bit pattern of value * 2^-24 as fp16), |v| <= 1023
    const uint32_t neg = (v & 0x80008000u) >> 15, m = neg * 0xFFFFu;
    return (add2(v ^ m, neg) & 0x7FFF7FFFu) | (m & 0x80008000u);
  }
  static __device__ __forceinline__ uint32_t from_sm(uint32_t v)
  {
    const uint32_t neg = (v & 0x80008000u) >> 15, m = neg * 0xFFFFu;
    return add2((v & 0x7FFF7FFFu) ^ m, neg);
  }
  __device__ __forceinline__ void cascade(uint32_t x)
  {
    const uint32_t y1 = add2(x, d1), y2 = add2(y1, d2), y3 = add2(y2, d3), y4 = add2(y3, d4), y5 = add2(y4, d5), y6 = add2(y5, d6);
    d1 = x; d2 = y1; d3 = y2; d4 = y3; d5 = y4; d6 = y5;
    o2 = o1;
    o1 = y6;
  }
  // the three trackers and the hit state, HBM form <-> register form (shared by the policies derived from this one)
  __device__ __forceinline__ void load_trackers(const uint32_t* st, uint32_t lane)
  {
    Mq = add2(~st[SV_MEDIAN * 32 + lane], 0x40024002u);       // ~m = -m - 1
    Q25n = add2(~st[SV_Q25 * 32 + lane], 0x40024002u);
#if SWTPG_FIR_INT
    A = add2(st[SV_ACCUM * 32 + lane], 0x000A000Au);
    Q75p = add2(~st[SV_Q75 * 32 + lane], 0x40024002u);
    A25 = add2(st[SV_A25 * 32 + lane], 0x000A000Au);
    A75 = add2(st[SV_A75 * 32 + lane], 0x000B000Bu);
#else
    A = to_sm(add2(st[SV_ACCUM * 32 + lane], 0xFFFFFFFFu));
    Q75p = add2(st[SV_Q75 * 32 + lane], 0x00020002u);
#if SWTPG_FIR_MERGED
    A25 = to_sm(add2(st[SV_A25 * 32 + lane], 0xFFFFFFFFu));
    A75 = to_sm(~st[SV_A75 * 32 + lane]);                      // ~a = -a - 1
#else
    A25 = to_sm(st[SV_A25 * 32 + lane]);
    A75 = to_sm(st[SV_A75 * 32 + lane]);
#endif
#endif
    prev = st[SV_PREV * 32 + lane];
    C = st[SV_CHARGE * 32 + lane];
    Tn = neg2(st[SV_TOVER * 32 + lane]);
  }
  __device__ __forceinline__ void store_trackers(uint32_t* st, uint32_t lane) const
  {
    st[SV_MEDIAN * 32 + lane] = median();
    st[SV_Q25 * 32 + lane] = add2(~Q25n, 0x40024002u);
#if SWTPG_FIR_INT
    st[SV_ACCUM * 32 + lane] = add2(A, 0xFFF6FFF6u);
    st[SV_Q75 * 32 + lane] = add2(~Q75p, 0x40024002u);
    st[SV_A25 * 32 + lane] = add2(A25, 0xFFF6FFF6u);
    st[SV_A75 * 32 + lane] = add2(A75, 0xFFF5FFF5u);
#else
    st[SV_ACCUM * 32 + lane] = add2(from_sm(A), 0x00010001u);
    st[SV_Q75 * 32 + lane] = add2(Q75p, 0xFFFEFFFEu);
#if SWTPG_FIR_MERGED
    st[SV_A25 * 32 + lane] = add2(from_sm(A25), 0x00010001u);
    st[SV_A75 * 32 + lane] = ~from_sm(A75);
#else
    st[SV_A25 * 32 + lane] = from_sm(A25);
    st[SV_A75 * 32 + lane] = from_sm(A75);
#endif
#endif
    st[SV_PREV * 32 + lane] = prev;
    st[SV_CHARGE * 32 + lane] = C;
    st[SV_TOVER * 32 + lane] = neg2(Tn);
  }
  __device__ __forceinline__ void load(const uint32_t* st, uint32_t lane, uint32_t flags)
  {
    load_trackers(st, lane);
    // ring -> cascade: replay the ring's 8 samples, oldest first, through an empty cascade (the state only depends on them)
    const uint32_t k = (flags >> 8) & 7u; // next slot to be written = oldest sample
    kphase0 = k;
    d1 = d2 = d3 = d4 = d5 = d6 = o1 = o2 = 0u;
    for (uint32_t i = 0; i < 8; ++i)
      cascade(st[(SV_RING0 + ((k + i) & 7u)) * 32 + lane]);
  }
  __device__ __forceinline__ uint32_t median() const { return add2(~Mq, 0x40024002u); } // 16385 - Mq
  __device__ __forceinline__ void store(uint32_t* st, uint32_t lane, uint32_t k_end) const
  {
    store_trackers(st, lane);
    // cascade -> ring, by running the cascade backwards: with Y[j] = y_j one tick ago (y_0 = the sample itself),
    // y_{j-1}(two ticks ago) = Y[j] - Y[j-1]; every step back in time loses the top stage, and y_0 is the ring entry.
    uint32_t Y[7] = { d1, d2, d3, d4, d5, d6, o1 };
    int top = 6;             // Y[0..top] valid
    uint32_t known_top = o2; // y_6 two ticks ago
#pragma unroll
    for (uint32_t i = 0; i < 8; ++i) {
      st[(SV_RING0 + ((k_end + 7u - i) & 7u)) * 32 + lane] = Y[0]; // sample i+1 ticks ago
      uint32_t Z[7] = {};
#pragma unroll
      for (int jj = 1; jj <= 6; ++jj)
        if (jj <= top)
          Z[jj - 1] = add2(Y[jj], neg2(Y[jj - 1]));
      if (i == 0) { // the column two ticks ago still has its top entry (o2)
        Z[6] = known_top;
      } else {
        top -= 1;
      }
#pragma unroll
      for (int jj = 0; jj < 7; ++jj)
        Y[jj] = Z[jj];
    }
  }
  uint32_t kphase0;
  __device__ __forceinline__ uint32_t phase_after(uint32_t ticks) const { return (kphase0 + ticks) & 7u; }
  // setState: pedestal = first sample, quartiles +-20 (wib2/tpg/ProcessingInfo.hpp:101-141)
  __device__ __forceinline__ void seed(uint32_t Sb) // Sb = ped - 16384
  {
    Mq = add2(~Sb, 0x00020002u);               // 16385 - ped = 1 - Sb
    Q25n = add2(~Sb, 0x00160016u);             // 16385 - (ped - 20)
#if SWTPG_FIR_INT
    Q75p = add2(~Sb, 0xFFEEFFEEu);             // 16385 - (ped + 20)
#else
    Q75p = add2(Sb, 0x40164016u);              // (ped + 20) + 2
#endif
  }

  // One frugal quartile step on the halves `en` (1.0 / 0.0 per half); c = sign(...) + 1 in {0,1,2} as the sign test delivers it.
  // Returns the step as two bit patterns 0 / 1 (up, down); the caller adds them to its register with the sign its form needs.
  static __device__ __forceinline__ void quartile_acc(uint32_t& Aq, uint32_t sgn, uint32_t en, uint32_t& up, uint32_t& dn)
  {
    const uint32_t T = hfma2_bits(en, sgn, Aq);                       // acc (+= sign on the enabled halves)
    up = hfma2_sat_bits(T, kOne, kNegL);                              // bit pattern 1 where acc reached  L+1
    dn = hfma2_sat_bits(T, kNegOne, kNegL);                           // ...                             -(L+1)
    Aq = hfma2_bits(ne2_abs_one(T, kUp), T, 0u);                      // reset where stepped
  }
  // The three frugal trackers of one tick (Sb = raw - 16384). Returns 16385 - median (integer form) or s' + 1 = raw - median + 1
  // (fp16 forms), updated median: see sp1_of / clamped_sample.
  // fp16 forms — ALU pipe: three sign tests, one complement, three IADD3, the final add. Everything else runs on the fp16 pipe.
  //
  // Integer form (SWTPG_FIR_INT, round 2). A frugal accumulator stays in [-10, 10] between ticks, so biased by +10 it is a small
  // NON-NEGATIVE number in each half, and so are the sign codes {0,1,2}: sums of such values are plain 32-bit adds (IADD3 issues on
  // either pipe at twice the rate of the packed-16x2 and fp16x2 instructions) because nothing carries between the halves. One
  // tracker step is
  //     U = B + sign + 1                      acc' + 11 in [0, 22]                                         IADD3
  //     c = clamp(U - 1, 0, 20)               acc' + 10 without the two step values                        VIADDMNMX.RELU
  //     D = U - 1 - c                         the step, -1 / 0 / +1 per half, as ONE 32-bit integer        IADD3
  //     B = c - 10 D  (or U - 11 D)           back to 10 (11) where it stepped                             IMAD
  //     Q = Q - D                             16385 - quantile follows                                     IADD3
  // D is a "true integer" d_hi * 65536 + d_lo whose low half may be negative (its bit pattern then borrows from the high half);
  // that is harmless because only LINEAR 32-bit operations consume it and every result they produce has non-negative halves
  // again, for which the 32-bit value and the packed value coincide. The quartile enables (:119-121: sample below / above the OLD
  // median) cost no select: the clamp's upper bound is 2 on the enabled halves and 0 elsewhere, which turns the sign code into 0
  // there, and the missing +1 comes from n1 (q25, "+10" form) or is taken back out by g1 (q75, "+11" form).
  // Per tick: 8 ALU-pipe instructions (sign tests, clamps, enables), 3-4 on the FMA pipe, 10 adds that go to whichever is free
  // — against 13 / 8 / 2 for the merged fp16 form and 9 / 14 / 3 for the separate one.
  __device__ __forceinline__ uint32_t track(uint32_t Sb)
  {
    const uint32_t sg1 = addclamp2(Sb, Mq, 0x00020002u);              // sign(raw - median) + 1, OLD median   (:108-117)
#if SWTPG_FIR_INT
    // q25 on the halves below the OLD median (:119)
    const uint32_t n1 = min2(sg1, 0x00010001u);                       // 0 where below, 1 elsewhere
    const uint32_t s25 = addclamp2(Sb, Q25n, n1 * (kM1 + kM1) + 0x00020002u); // bound 2 - 2 n1 (IMAD): sign(raw - q25) + 1 where enabled, else 0
    const uint32_t U25 = SWTPG_FIR_U3 ? fma_add(n1, fma_add(s25, A25, kP1), kP1) : A25 + s25 + n1; // acc25' + 11
    const uint32_t c25 = addclamp2(U25, kM1, 0x00140014u);
    const uint32_t D25 = U25 - c25 - 0x00010001u;
    A25 = c25 - 10u * D25;
    Q25n = fma_sub(Q25n, D25, kM1);
    // q75 on the halves above it (:121)
    const uint32_t g1 = __viaddmax_s16x2_relu(sg1, kM1, 0u);          // 1 where above, 0 elsewhere
    const uint32_t s75 = addclamp2(Sb, Q75p, fma_add(g1, g1, kP1));  // bound 2 g1: sign(raw - q75) + 1 where enabled, else 0
    const uint32_t U75 = SWTPG_FIR_U3 ? fma_sub(fma_add(s75, A75, kP1), g1, kM1) : A75 + s75 - g1; // acc75' + 11
    const uint32_t c75 = addclamp2(U75, kM1, 0x00140014u);
    const uint32_t D75 = U75 - c75 - 0x00010001u;
    A75 = U75 - 11u * D75;
    Q75p = fma_sub(Q75p, D75, kM1);
    // median (:125)
    const uint32_t U = fma_add(sg1, A, kP1);
    const uint32_t cm = addclamp2(U, kM1, 0x00140014u);
    const uint32_t D = U - cm - 0x00010001u;
    A = cm - 10u * D;
    Mq = fma_sub(Mq, D, kM1);
    return Mq;
#else
    const uint32_t ltf = eq2_one(sg1, 0u), gtf = eq2_one(sg1, 0x00020002u);
#if SWTPG_FIR_MERGED
    // q25 moves only on the halves below the OLD median and q75 only on the halves above it (:119-121): per half at most one of
    // them does anything, so ONE frugal step runs on per-half SELECTED operands — (Sb, 16385 - q25, acc25) below, (~raw, q75 + 2,
    // -acc75) elsewhere: with the sample and the accumulator negated the q75 tracker obeys the same "register += down - up"
    // rule as the q25 one — and is written back through the two masks; where the sample equals the median it is dropped.
    (void)ltf;
    (void)gtf;
    const uint32_t lt = eq2_mask(sg1, 0u), gt = eq2_mask(sg1, 0x00020002u);
    const uint32_t a = (lt & Sb) | (~lt & (~Sb | 0xC000C000u));       // raw - 16384 below; ~raw = -raw - 1 elsewhere (one LOP3)
    const uint32_t Qs = (lt & Q25n) | (~lt & Q75p), As = (lt & A25) | (~lt & A75);
    const uint32_t Tq = hadd2_bits(As, addclamp2(a, Qs, 0x00020002u)); // acc25 | -acc75 after this sample
    const uint32_t uq = hfma2_sat_bits(Tq, kOne, kNegL), dq = hfma2_sat_bits(Tq, kNegOne, kNegL);
    const uint32_t An = hfma2_bits(ne2_abs_one(Tq, kUp), Tq, kNegTiny); // (stepped ? 0 : acc) - 1
    const uint32_t Qn = Qs + dq - uq;
    Q25n = (lt & Qn) | (~lt & Q25n);
    A25 = (lt & An) | (~lt & A25);
    Q75p = (gt & Qn) | (~gt & Q75p);
    A75 = (gt & An) | (~gt & A75);
#else
    uint32_t up, dn;
    // q25 on the halves below the median (:119): sign(raw - q25) from Sb + (16385 - q25)
    quartile_acc(A25, hadd2_bits(addclamp2(Sb, Q25n, 0x00020002u), kNegTiny), ltf, up, dn);
    Q25n = Q25n + dn - up;                                            // 16385 - q25 steps down when q25 steps up
    // q75 on the halves above (:121): 1 - (sign(q75 - raw) + 1) from ~raw + (q75 + 2), ~raw = -raw - 1 = ~Sb | 0xC000 per half
    quartile_acc(A75, hfma2_bits(addclamp2(~Sb | 0xC000C000u, Q75p, 0x00020002u), kNegOne, kTiny), gtf, up, dn);
    Q75p = Q75p + up - dn;
#endif
    // median (:125), as PackedSimpleT::pedestal_step with L = 10
    const uint32_t T = hadd2_bits(A, sg1);
    const uint32_t up1 = hfma2_sat_bits(T, kOne, kNegL), dn1 = hfma2_sat_bits(T, kNegOne, kNegL);
    A = hfma2_bits(ne2_abs_one(T, kUp), T, kNegTiny);
    Mq = Mq + dn1 - up1;
    return add2(Sb, Mq);
#endif
  }
  // min(q75 - q25, sigmaMax) + 3 (:131-134) from a pair of quartile registers (Q75p, Q25n)
  __device__ __forceinline__ uint32_t sig3_of(uint32_t q75p, uint32_t q25n) const
  {
#if SWTPG_FIR_INT
    return addmin2(add2(~q75p, 0x00040004u), q25n, sig3max);          // -(16385 - q75) + 3 = ~Q75p + 4 per half
#else
    return addmin2(add2(q75p, 0xC000C000u), q25n, sig3max);
#endif
  }
  // One tick: returns the filter output of this tick.
  // s' + 1 = raw - median + 1 from what track() returns
  static __device__ __forceinline__ uint32_t sp1_of(uint32_t Sb, uint32_t tr)
  {
#if SWTPG_FIR_INT
    return add2(Sb, tr);
#else
    return tr;
#endif
  }
  // min(raw - median, adcMax) (:128,142) from what track() returns
  __device__ __forceinline__ uint32_t clamped_sample(uint32_t Sb, uint32_t tr) const
  {
#if SWTPG_FIR_INT
    return addmin2(Sb, fma_add(tr, 0xFFFEFFFFu, kP1), xmax);         // tr = 16385 - median >= 2: the -1 is a plain 32-bit add
#else
    return addmin2(tr, 0xFFFFFFFFu, xmax);                            // tr = raw - median + 1
#endif
  }
  __device__ __forceinline__ uint32_t tick(uint32_t Sb)
  {
    const uint32_t x = clamped_sample(Sb, track(Sb));
    const uint32_t filt = o2;                                         // window = samples t-8 .. t-2          (:160-201)
    cascade(x);
    return filt;
  }
  // sigma * multiplier * threshold for sigma >= 0: (sig3 - 3) * K per half, as one 32-bit multiply-add — no carry crosses
  // the halves because each half of sig3 * K is >= 3K and < 2^16 + 3K... the host checks (sigmaMax + 3) * K < 2^16.
  __device__ __forceinline__ uint32_t threshold(uint32_t sig3) const { return sig3 * K + Kneg3; }

  template<bool EXACT>
  __device__ __forceinline__ void hit_update(uint32_t filt, uint32_t thr, const TickCtx& ctx, int t)
  {
    uint32_t over;
    if constexpr (EXACT) // any threshold value: plain signed compares
      over = (lo16s(filt) > lo16s(thr) ? 0xFFFFu : 0u) | (hi16s(filt) > hi16s(thr) ? 0xFFFF0000u : 0u);
    else
      over = gt2_mask_bf16(max2(filt, 0u), thr);          // 0 <= thr <= 32640                        (:208)
    const uint32_t left = prev & ~over;                   //                                           (:210)
    uint32_t add;
    if constexpr (EXACT)
      add = pack2((lo16s(filt & over)) >> shift, (hi16s(filt & over)) >> shift);
    else
      add = ((filt & over) >> shift) & shmask;            // over => filt > thr >= 0: arithmetic == logical shift (:221-223)
    if constexpr (EXACT)
      C = pack2(sat16(lo16s(C) + lo16s(add)), sat16(hi16s(C) + hi16s(add)));
    else
      C = minu2(add2(C, add), 0x7FFF7FFFu);               // adds_epi16 of two non-negative halves
    Tn = addmax2(Tn, over, 0x80018001u);                  // tover = adds(tover, 1)                    (:236-237)
    prev = over;
    if (left != 0u) {                                     //                                           (:251-281)
      // accepted iff hit_charge != 0 (src/wib2/WIB2FrameProcessor.cpp:429): decided in flush on the masked charge
      ctx.stage->push(HitStage::meta(ctx.chan0, ctx.unit, uint32_t(t)), C & left, neg2(Tn));
      C &= ~left;
      Tn &= ~left;
    }
  }

  // Quiet test and hit bookkeeping of one 4-tick group, given its filter outputs and the quartile registers after each tick.
  template<int G, bool WIB2_UNITS>
  __device__ __forceinline__ void finish_group(const uint32_t (&filt)[G], const uint32_t (&q75v)[G], const uint32_t (&q25v)[G], const TickCtx& ctx,
                                               int t0)
  {
    // Conservative quiet test: the largest filter output of the group against the threshold of a LOWER BOUND of its sigmas.
    // At most one quartile moves per tick, by one, so every sigma of the group is >= sigma(last tick) - 3.
    const uint32_t mx = __vimax3_s16x2(__vimax3_s16x2(filt[0], filt[1], filt[2]), filt[3], 0u);
    const uint32_t lo3 = add2(sig3_of(q75v[G - 1], q25v[G - 1]), 0xFFFDFFFDu); // (bound of min(sigma, sigmaMax)) + 3
    uint32_t neg = addmin2(lo3, 0xFFFDFFFDu, 0u);        // non-zero <=> the bound is negative: some sigma MAY be below zero
    if (ctx.p->debug_flags & 1u)
      neg = 0xFFFFFFFFu;
    const uint32_t busy = gt2_mask_bf16(mx, threshold(lo3)) | prev | neg;
    if (__builtin_expect(!__any_sync(0xFFFFFFFFu, busy != 0u), 1))
      return; // nothing but the trackers and the filter moves outside hits
    uint32_t sig3[G];
#pragma unroll
    for (int g = 0; g < G; ++g)
      sig3[g] = sig3_of(q75v[g], q25v[g]);
    const uint32_t smin = __vimin3_s16x2(__vimin3_s16x2(sig3[0], sig3[1], sig3[2]), sig3[3], sig3[3]);
    uint32_t really_neg = addmin2(smin, 0xFFFDFFFDu, 0u); // min(sigma, 0) of the group: non-zero <=> some sigma < 0
    if (ctx.p->debug_flags & 1u)
      really_neg = 0xFFFFFFFFu;
    if (__any_sync(0xFFFFFFFFu, really_neg != 0u)) { // rare: carries between the positions of a 64-bit lane (H7)
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const uint32_t th = iqr_threshold_exact(add2(sig3[g], 0xFFFDFFFDu), (ctx.chan0 >> 1) & 31u, int(mult), thr_cfg);
        hit_update<true>(filt[g], th, ctx, t0 + g);
      }
    } else {
#pragma unroll
      for (int g = 0; g < G; ++g)
        hit_update<false>(filt[g], threshold(sig3[g]), ctx, t0 + g);
    }
    if (ctx.stage->must_flush())
      flush<WIB2_UNITS>(*ctx.stage, ctx.p->sink, ctx.link_base, ctx.link, (ctx.chan0 >> 1) & 31u, false); // whole 32-pair rounds only
  }

  template<int G, bool DUMP, int ROW_WORDS = 28, bool WIB2_UNITS = false>
  __device__ __forceinline__ void group(const uint32_t* rows, const PairPos& pp, const TickCtx& ctx, int t0, uint32_t* ped_out,
                                        uint32_t* wav_out, bool /*more*/)
  {
    static_assert(G == 4, "trees below are written for 4 ticks");
    uint32_t filt[G], q75v[G], q25v[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      filt[g] = tick(extract_pair_biased(rows + g * ROW_WORDS, pp));
      q75v[g] = Q75p;
      q25v[g] = Q25n;
      if constexpr (DUMP) {
        ped_out[g] = median();
        wav_out[g] = filt[g];
      }
    }
    finish_group<G, WIB2_UNITS>(filt, q75v, q25v, ctx, t0);
  }
};

// =====================================================================================================================
// Packed FIR + IQR for ARBITRARY taps (any int16 taps[0..6], e.g. firwin_int at another multiplier, or a user-supplied filter):
// same trackers, threshold and hit bookkeeping as PackedFirIqr, but the filter is the reference's multiply-add chain
//   filt = sum_{j<7} taps[j] * prev_samp[(j + k) % 8]        (mullo_epi16 / add_epi16, mod 2^16; ProcessAVX2FIR.hpp:160-201)
// on a window of the last 8 clamped samples held in registers, oldest first (w[0] = s(t-8) ... w[7] = s(t-1); the ring phase
// only matters when the state is loaded and stored). Products are 32-bit IMADs whose low 16 bits are the wrapped result, so
// the packed register itself is the low-half operand; a half-swapped copy of every window entry serves the high half.
// 14 IMAD + 2 PRMT per tick instead of the binomial cascade's 6 adds.
// =====================================================================================================================
struct PackedFirIqrAnyTaps : PackedFirIqr
{
  static constexpr int kQuadCtasPerSm = 4; // CTA form, 16 consumer warps per SM: 6 % faster than one warp per CTA
  uint32_t w[8], ws[8]; // window and its half-swapped twin
  int tap[7];

  __device__ __forceinline__ void configure(const KernelParams& p)
  {
    PackedFirIqr::configure(p);
#pragma unroll
    for (int j = 0; j < 7; ++j)
      tap[j] = p.taps[j];
  }
  static __device__ __forceinline__ uint32_t swap_halves(uint32_t v) { return __byte_perm(v, v, 0x1032); }
  __device__ __forceinline__ void load(const uint32_t* st, uint32_t lane, uint32_t flags)
  {
    load_trackers(st, lane);
    const uint32_t k = (flags >> 8) & 7u; // next slot to be written = oldest sample
    kphase0 = k;
#pragma unroll
    for (uint32_t i = 0; i < 8; ++i) {
      w[i] = st[(SV_RING0 + ((k + i) & 7u)) * 32 + lane];
      ws[i] = swap_halves(w[i]);
    }
  }
  __device__ __forceinline__ void store(uint32_t* st, uint32_t lane, uint32_t k_end) const
  {
    store_trackers(st, lane);
#pragma unroll
    for (uint32_t i = 0; i < 8; ++i)
      st[(SV_RING0 + ((k_end + i) & 7u)) * 32 + lane] = w[i]; // w[0] is the oldest = the slot written next
  }
  __device__ __forceinline__ uint32_t tick(uint32_t Sb)
  {
    const uint32_t x = clamped_sample(Sb, track(Sb));
    int lo = 0, hi = 0;
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      lo += tap[j] * int(w[j]);   // low 16 bits: taps[j] * (low half), any upper bits are discarded below
      hi += tap[j] * int(ws[j]);
    }
    const uint32_t filt = __byte_perm(uint32_t(lo), uint32_t(hi), 0x5410);
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      w[j] = w[j + 1];
      ws[j] = ws[j + 1];
    }
    w[7] = x;
    ws[7] = swap_halves(x);
    return filt;
  }

  template<int G, bool DUMP, int ROW_WORDS = 28, bool WIB2_UNITS = false>
  __device__ __forceinline__ void group(const uint32_t* rows, const PairPos& pp, const TickCtx& ctx, int t0, uint32_t* ped_out,
                                        uint32_t* wav_out, bool /*more*/)
  {
    static_assert(G == 4, "trees below are written for 4 ticks");
    uint32_t filt[G], q75v[G], q25v[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      filt[g] = tick(extract_pair_biased(rows + g * ROW_WORDS, pp));
      q75v[g] = Q75p;
      q25v[g] = Q25n;
      if constexpr (DUMP) {
        ped_out[g] = median();
        wav_out[g] = filt[g];
      }
    }
    finish_group<G, WIB2_UNITS>(filt, q75v, q25v, ctx, t0);
  }
};

// =====================================================================================================================
// Packed fast path: WIB2 AbsRS (wib2/tpg/ProcessRSAVX2.hpp:24-330) — the IQR threshold of the FIR finder applied to an absolute
// running sum. Per tick and channel:
//   quartile / median trackers exactly as PackedFirIqr (every limit 10, :100-126);  s' = raw - median
//   RS = mulhrs(RS * 8 + |s'| * 5, 32768/10)      literal factors (:29-33,:137-150), mullo / add mod 2^16
//   RS -= median_RS after a frugal update of median_RS with RS (:152-159)
//   sigma = min(q75 - q25, sigmaMax), sigmaMax = 2^15 / (multiplier * threshold) (:36);  over = RS > sigma * threshold (:198)
//   charge = adds(charge, (over ? adds(RS, median_RS) : 0) >> tap_exponent) (:210-213);  tover = adds(tover, over)
// |RS before subtraction| <= 3277 (mulhrs of a 16-bit value by 3276), so adds(RS, median_RS) is that value again and never
// saturates. The threshold is one packed IMAD while every sigma of the warp is >= 0 (as in PackedFirIqr); a group that sees a
// negative sigma recomputes it with the reference's 4 x int64-lane semantics. Validity (checked by the host):
// 1 <= tap_exponent <= 10, threshold >= 1, (sigmaMax + 3) * threshold < 2^16.
// =====================================================================================================================
struct PackedRsIqrWib2 : PackedFirIqr
{
  static constexpr int kGroupUnroll = 1;
  static constexpr int kWib2MinCtas = 1;
  uint32_t RS1, MRq, AR; // RS + 1 (after median subtraction); 1 - median_RS; (acc_RS - 1) as fp16 subnormal

  __device__ __forceinline__ void configure(const KernelParams& p)
  {
    PackedFirIqr::configure(p);
    const uint32_t sigma_max = (1u << 15) / (mult * p.threshold); // ProcessRSAVX2.hpp:36
    sig3max = (sigma_max + 3u) * 0x00010001u;
    K = p.threshold;                                             // sigma * info.threshold: no multiplier (:198)
    Kneg3 = 0u - 3u * K * 0x00010001u;
  }
  __device__ __forceinline__ void load(const uint32_t* st, uint32_t lane, uint32_t)
  {
    load_trackers(st, lane);
    RS1 = add2(st[SV_RS * 32 + lane], 0x00010001u);
    MRq = add2(~st[SV_MED_RS * 32 + lane], 0x00020002u);
    AR = to_sm(add2(st[SV_ACC_RS * 32 + lane], 0xFFFFFFFFu));
    kphase0 = 0;
  }
  __device__ __forceinline__ void store(uint32_t* st, uint32_t lane, uint32_t) const
  {
    store_trackers(st, lane);
    st[SV_RS * 32 + lane] = add2(RS1, 0xFFFFFFFFu);
    st[SV_MED_RS * 32 + lane] = add2(~MRq, 0x00020002u);
    st[SV_ACC_RS * 32 + lane] = add2(from_sm(AR), 0x00010001u);
  }
  __device__ __forceinline__ uint32_t phase_after(uint32_t) const { return 0; }

  // One tick: sp1 = s' + 1. Returns RS - median_RS + 1 and, in `raw_rs`, the running sum before the subtraction.
  __device__ __forceinline__ uint32_t rs_step(uint32_t sp1, uint32_t& raw_rs)
  {
    const uint32_t x = add2(sp1, 0xFFFFFFFFu);
    const uint32_t ax = max2(x, neg2(x));                                     // |s'|
    // low 16 bits per half: RS * 8 + |s'| * 5, with RS = RS1 - 1 (the packed registers serve as low-half operands)
    const uint32_t lo = uint32_t(int(RS1) * 8 + (int(ax) * 5 - 8));
    const uint32_t hi = uint32_t(int(RS1 >> 16) * 8 + (int(ax >> 16) * 5 - 8));
    const int r_lo = PackedRsWibEth<false>::mulhrs_low(lo), r_hi = PackedRsWibEth<false>::mulhrs_low(hi);
    raw_rs = __byte_perm(uint32_t(r_lo), uint32_t(r_hi), 0x5410);
    // frugal update of median_RS with the new running sum (limit 10), then subtract
    const uint32_t sg1 = addclamp2(raw_rs, MRq, 0x00020002u);
    const uint32_t T = hadd2_bits(AR, sg1);
    const uint32_t upm = eq2_mask(T, kUp);
    const uint32_t dn1 = hfma2_sat_bits(T, kNegOne, kNegL);
    AR = hfma2_bits(ne2_abs_one(T, kUp), T, kNegTiny);
    MRq = add2(add2(MRq, upm), dn1);
    RS1 = add2(raw_rs, MRq);
    return RS1;
  }

  template<bool EXACT>
  __device__ __forceinline__ void hit_update(uint32_t lv1, uint32_t raw_rs, uint32_t thr, const TickCtx& ctx, int t)
  {
    const uint32_t lv = add2(lv1, 0xFFFFFFFFu);
    uint32_t over;
    if constexpr (EXACT)
      over = (lo16s(lv) > lo16s(thr) ? 0xFFFFu : 0u) | (hi16s(lv) > hi16s(thr) ? 0xFFFF0000u : 0u);
    else
      over = gt2_mask_bf16(max2(lv, 0u), thr);            // 0 <= thr <= 32640                        (:198-200)
    const uint32_t left = prev & ~over;
    // (over ? adds(RS, median_RS) : 0) >> tap_exponent, arithmetic per half (the sum may be negative): offset-binary trick
    const uint32_t u = (add2(raw_rs & over, 0x80008000u) >> shift) & shmask;
    const uint32_t bias = (0x8000u >> shift) * 0x00010001u;
    const uint32_t add = add2(u, neg2(bias));
    const uint32_t Cw = add2(C, add);                     // adds_epi16: wrapping add + repair of the rare overflowing halves
    const uint32_t ovf = ~(C ^ add) & (C ^ Cw) & 0x80008000u;
    if (__builtin_expect(ovf != 0u, 0))
      C = pack2(sat16(lo16s(C) + lo16s(add)), sat16(hi16s(C) + hi16s(add)));
    else
      C = Cw;
    Tn = addmax2(Tn, over, 0x80018001u);                  // tover = adds(tover, 1)
    prev = over;
    if (left != 0u) {
      // accepted iff hit_charge != 0 (src/wib2/WIB2FrameProcessor.cpp:429): decided in flush on the masked charge
      ctx.stage->push(HitStage::meta(ctx.chan0, ctx.unit, uint32_t(t)), C & left, neg2(Tn));
      C &= ~left;
      Tn &= ~left;
    }
  }

  template<int G, bool DUMP, int ROW_WORDS = 28, bool WIB2_UNITS = false>
  __device__ __forceinline__ void group(const uint32_t* rows, const PairPos& pp, const TickCtx& ctx, int t0, uint32_t* ped_out,
                                        uint32_t* wav_out, bool /*more*/)
  {
    static_assert(G == 4, "trees below are written for 4 ticks");
    uint32_t lv[G], rs[G], sig3[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const uint32_t Sb = extract_pair_biased(rows + g * ROW_WORDS, pp);
      lv[g] = rs_step(sp1_of(Sb, track(Sb)), rs[g]);
      sig3[g] = sig3_of(Q75p, Q25n);
      if constexpr (DUMP) {
        ped_out[g] = median();
        wav_out[g] = add2(lv[g], 0xFFFFFFFFu);
      }
    }
    // conservative quiet test: largest RS of the group (biased by one: still conservative) vs the threshold of its smallest sigma
    const uint32_t mx = __vimax3_s16x2(__vimax3_s16x2(lv[0], lv[1], lv[2]), lv[3], 0u);
    const uint32_t smin = __vimin3_s16x2(__vimin3_s16x2(sig3[0], sig3[1], sig3[2]), sig3[3], sig3[3]);
    uint32_t neg = addmin2(smin, 0xFFFDFFFDu, 0u); // min(sigma, 0) of the group: non-zero <=> some sigma < 0
    if (ctx.p->debug_flags & 1u)
      neg = 0xFFFFFFFFu;
    const uint32_t busy = gt2_mask_bf16(mx, threshold(smin)) | prev | neg;
    if (__builtin_expect(!__any_sync(0xFFFFFFFFu, busy != 0u), 1))
      return; // nothing but the trackers and the running sum moves outside hits
    if (__any_sync(0xFFFFFFFFu, neg != 0u)) { // rare: carries between the positions of a 64-bit lane (H7)
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const uint32_t th = iqr_threshold_exact(add2(sig3[g], 0xFFFDFFFDu), (ctx.chan0 >> 1) & 31u, 1, thr_cfg);
        hit_update<true>(lv[g], rs[g], th, ctx, t0 + g);
      }
    } else {
#pragma unroll
      for (int g = 0; g < G; ++g)
        hit_update<false>(lv[g], rs[g], threshold(sig3[g]), ctx, t0 + g);
    }
    if (ctx.stage->must_flush())
      flush<WIB2_UNITS>(*ctx.stage, ctx.p->sink, ctx.link_base, ctx.link, (ctx.chan0 >> 1) & 31u, false); // whole 32-pair rounds only
  }
};

// =====================================================================================================================
// WIBEth kernel: persistent warps, one link (64 channels) at a time per warp, per-warp TMA ring of CHUNK_TICKS-tick stages.
// Warp w of the grid starts on link w; every further link is CLAIMED from a global cursor when the warp's producer runs
// out of chunks, so links go to whichever warp is free first: a warp that shares its SM sub-partition with one more
// neighbour, or that drew longer links in a ragged batch, simply takes fewer of them. The ring is refilled by lane 0 from
// a producer cursor that runs NSTAGE chunks ahead of the consumer ACROSS link boundaries, so a warp never drains its
// pipeline between links; the producer hands the links it chose to the consumer through a small warp-uniform FIFO.
// =====================================================================================================================
constexpr int kWibEthRowBytes = 112;

// Dynamic shared memory of one CTA: [stages | mbarriers | hit staging (16 B + 4 B per record) | hit counters | link FIFOs].
template<int WARPS, int NSTAGE, int CHUNK_TICKS>
struct WibEthSmem
{
  static constexpr size_t align16(size_t x) { return (x + 15) & ~size_t(15); }
  static constexpr size_t bars = align16(size_t(WARPS) * NSTAGE * kWibEthRowBytes * CHUNK_TICKS);
  static constexpr size_t hits = align16(bars + size_t(WARPS) * NSTAGE * 8);
  static constexpr size_t aux = hits + size_t(WARPS) * HitStage::kCap * 16;
  static constexpr size_t counts = aux + size_t(WARPS) * HitStage::kCap * 4;
  static constexpr size_t fifo = align16(counts + size_t(WARPS) * 4);
  static constexpr size_t total = fifo + size_t(WARPS) * 32; // 8-entry link FIFO per warp
};

// Whole warp: wait until *counter == want (acquire: what the publishing warp stored before its release is visible afterwards).
__device__ __noinline__ void
wait_for_slices(const uint32_t* counter, uint32_t want)
{
  uint32_t seen;
  for (;;) {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    if (__all_sync(0xFFFFFFFFu, seen == want))
      break;
    __nanosleep(200);
  }
}

template<class Algo, int WARPS, int NSTAGE, int CHUNK_TICKS, bool DUMP, int MIN_CTAS = 1>
__global__ void __launch_bounds__(WARPS * 32, MIN_CTAS)
wibeth_kernel(const KernelParams p)
{
  static_assert(64 % CHUNK_TICKS == 0, "chunk must divide the frame");
  constexpr int kChunkBytes = kWibEthRowBytes * CHUNK_TICKS;
  constexpr int kChunksPerUnit = 64 / CHUNK_TICKS;
  extern __shared__ __align__(128) uint8_t smem[];

  // warp index through a shuffle: tells the compiler it is warp-uniform, so the ring bookkeeping runs on the uniform datapath
  const uint32_t warp = __shfl_sync(0xFFFFFFFFu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31u;
  const uint32_t warps_total = gridDim.x * WARPS;
  const uint32_t first_link = blockIdx.x * WARPS + warp; // this warp's first work item
  const uint32_t n_items = p.n_links << p.parts_log2;
  if (first_link >= n_items)
    return; // warps are fully independent: no block-level barrier anywhere below

  using L = WibEthSmem<WARPS, NSTAGE, CHUNK_TICKS>;
  uint8_t* stages = smem + size_t(warp) * NSTAGE * kChunkBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::bars) + warp * NSTAGE;
  HitStage hits;
  hits.buf = reinterpret_cast<uint4*>(smem + L::hits) + size_t(warp) * HitStage::kCap;
  hits.aux = reinterpret_cast<uint32_t*>(smem + L::aux) + size_t(warp) * HitStage::kCap;
  hits.cnt = reinterpret_cast<uint32_t*>(smem + L::counts) + warp;

  auto units_of = [&](uint32_t link) -> uint32_t { return p.n_units ? p.n_units[link] : p.units_stride; };
  auto base_of = [&](uint32_t link) -> const uint8_t* { return p.frames + size_t(link) * p.units_stride * SWTPG_WIBETH_FRAME_BYTES; };

  // Producer cursor: the next chunk to request is pr_src; pr_left chunks remain in link pr_link. Every lane keeps the same
  // (warp-uniform) copy, so the bookkeeping runs on the uniform datapath; only lane 0 talks to the mbarrier / copy engine
  // and to the global link cursor.
  // FIFO of links the producer has started and the consumer has not: the producer is at most NSTAGE chunks ahead and every
  // link it starts has >= kChunksPerUnit chunks, so at most ceil(NSTAGE / kChunksPerUnit) + 1 links are in flight.
  constexpr uint32_t kFifo = 8; // power of two; lives in shared memory so the tick loop carries no registers for it
  static_assert((NSTAGE + kChunksPerUnit - 1) / kChunksPerUnit + 1 <= int(kFifo), "link FIFO too short for this ring geometry");
  volatile uint32_t* fifo = reinterpret_cast<volatile uint32_t*>(smem + L::fifo) + warp * kFifo;
  uint32_t fifo_head = 0, fifo_n = 0;
  auto fifo_push = [&](uint32_t link) { // whole warp, converged; every lane holds the same `link`
    if (lane == 0)
      fifo[(fifo_head + fifo_n) & (kFifo - 1u)] = link;
    ++fifo_n;
    __syncwarp();
  };
  auto fifo_pop = [&]() -> uint32_t {
    const uint32_t link = fifo[fifo_head];
    fifo_head = (fifo_head + 1u) & (kFifo - 1u);
    --fifo_n;
    return link;
  };
  // Work items: item i = slice (i / n_links) of link (i % n_links); slice k of a link with n units is its units
  // [k n >> parts_log2, (k + 1) n >> parts_log2), or, with slice_geom, halving slices (what the launch's last round leaves idle is
  // at most one LAST slice, so small last slices shorten the tail without more state round trips per link). Items are claimed
  // in increasing order, so every first slice is handed out before any second one.
  auto slice_of = [&](uint32_t item, uint32_t& link, uint32_t& u0, uint32_t& u1) -> uint32_t { // returns the slice number
    uint32_t part = 0;
    link = item;
    while (link >= p.n_links) { // at most 7 rounds; warp-uniform
      link -= p.n_links;
      ++part;
    }
    const uint32_t n = units_of(link);
    if (p.slice_geom) { // halving slices: n/2, n/4, ... and the rest (slice k starts ceil(n / 2^k) units before the end)
      u0 = n - ((n + (1u << part) - 1u) >> part);
      u1 = part + 1u == (1u << p.parts_log2) ? n : n - ((n + (2u << part) - 1u) >> (part + 1u));
    } else {
      u0 = (part * n) >> p.parts_log2;
      u1 = ((part + 1u) * n) >> p.parts_log2;
    }
    return part;
  };
  // How many of the slices before `part` hold units (with fewer units than slices some are empty and never run). Equal slices:
  // below one unit per slice every non-empty one holds exactly one, so u0 of them; halving slices: slice k < last is empty
  // unless n > 2^k. The last slice is never empty in either scheme (it re-arms the link's counter).
  auto slices_before = [&](uint32_t link, uint32_t part, uint32_t u0) -> uint32_t {
    if (!p.slice_geom)
      return min(part, u0);
    const uint32_t n = units_of(link);
    return min(part, n > 1u ? 32u - uint32_t(__clz(int(n - 1u))) : 0u);
  };
  // same value in every lane -> uniform register, so that the producer's per-chunk bookkeeping stays on the uniform datapath
  auto uniform = [](uint32_t v) { return __reduce_max_sync(0xFFFFFFFFu, v); };
  bool pr_done = false; // the cursor has run past the last item
  uint32_t pr_left = 0, pr_in_unit = 0, pr_slot = 0;
  const uint8_t* pr_src = nullptr;
  auto pr_start = [&](uint32_t item) { // point the producer at an item (pr_left stays 0 if it is empty)
    uint32_t link, u0, u1;
    slice_of(item, link, u0, u1);
    pr_left = uniform((u1 - u0) * kChunksPerUnit);
    if (pr_left != 0) {
      fifo_push(item);
      pr_src = base_of(uniform(link)) + size_t(uniform(u0)) * SWTPG_WIBETH_FRAME_BYTES + 32;
      pr_in_unit = 0;
    }
  };
  auto claim = [&]() -> uint32_t { // next unclaimed item (items 0 .. warps_total-1 are the warps' first ones)
    uint32_t k = 0;
    if (lane == 0)
      k = atomicAdd(p.link_cursor, 1u);
    return warps_total + __reduce_add_sync(0xFFFFFFFFu, k); // REDUX: the result lands in a uniform register
  };
  pr_start(first_link);
  auto produce = [&]() { // request one more chunk, if any item is left for this warp (whole warp calls, converged)
    if (__builtin_expect(pr_left == 0, 0)) { // item exhausted (or empty): claim the next one that has data
      if (pr_done)
        return;
      do {
        const uint32_t item = claim();
        if (item >= n_items) {
          pr_done = true;
          return;
        }
        pr_start(item);
      } while (pr_left == 0);
    }
    if (SWTPG_ELECT ? elect_one() : lane == 0) {
      mbar_arrive_expect_tx(&bars[pr_slot], kChunkBytes);
      bulk_g2s(stages + pr_slot * kChunkBytes, pr_src, kChunkBytes, &bars[pr_slot]);
    }
    pr_slot = pr_slot + 1 == NSTAGE ? 0 : pr_slot + 1;
    pr_src += kChunkBytes;
    if (++pr_in_unit == kChunksPerUnit) { // skip the 32 header bytes of the next frame
      pr_in_unit = 0;
      pr_src += 32;
    }
    --pr_left;
  };

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s)
      mbar_init(&bars[s], 1);
    *hits.cnt = 0u;
    fence_mbar_init();
  }
  __syncwarp();
  for (int s = 0; s < NSTAGE; ++s)
    produce();

  Algo algo;
  algo.configure(p);
  const PairPos pp = pair_pos(lane);
  TickCtx ctx;
  ctx.p = &p;
  ctx.chan0 = 2 * lane;
  ctx.ts = 0;
  ctx.stage = &hits;

  uint32_t stg = 0, phase = 0; // consumer position in the ring and its mbarrier phase
  while (fifo_n != 0) { // the producer started this link NSTAGE chunks ago (or at start-up)
    uint32_t link, u0, n_units; // n_units: end of the slice
    const uint32_t part = slice_of(fifo_pop(), link, u0, n_units);
    const uint8_t* link_base = base_of(link);
    uint32_t* st = p.state + size_t(link) * kStateWordsPerGroup;
    // A later slice of a link continues from the state its predecessor stored: wait until the NON-EMPTY slices before it are
    // done (with fewer units than slices some are empty and never run: exactly min(part, u0) of the earlier ones are not). By
    // the time a second slice is claimed every first one was handed out a whole round ago, so this hardly ever spins; it cannot
    // dead-lock because every warp of the persistent grid is resident and items are taken in dependency order.
    if (part != 0u && u0 != 0u) {
      const uint32_t before = slices_before(link, part, u0);
      wait_for_slices(p.link_done + link, before); // out of line: the spin must not cost the tick loop registers
    }
    const uint32_t flags = *reinterpret_cast<volatile uint32_t*>(p.group_flags + link);
    algo.load(st, lane, flags);
    bool need_seed = !(flags & kFlagInitialized);
    ctx.link = link;
    ctx.link_base = link_base;

    for (uint32_t unit = u0; unit < n_units; ++unit) {
      ctx.tick_base = 0u; // 64 ticks per unit = 0 mod 8: the FIR ring phase at a unit's first tick is the slice's initial one
      ctx.unit = unit;
      if constexpr (!std::is_same<Algo, PackedSimpleWibEth>::value && !std::is_same<Algo, PackedSimpleWibEthPipe>::value) {
        // DAQEthHeader word 1 = timestamp (docs/README.md:81); the packed path reads it when it flushes hits
        ctx.ts = *reinterpret_cast<const unsigned long long*>(link_base + size_t(unit) * SWTPG_WIBETH_FRAME_BYTES + 8);
      }
#pragma unroll 1
      for (int t0 = 0; t0 < 64; t0 += CHUNK_TICKS) {
        mbar_wait(&bars[stg], phase);
        const uint32_t* rows = reinterpret_cast<const uint32_t*>(stages + stg * kChunkBytes);
        if (need_seed) {
          algo.seed(Algo::sample(rows, pp));
          need_seed = false;
        }
        constexpr int G = 4;
        static_assert(CHUNK_TICKS % G == 0, "group must divide the chunk");
        constexpr int kGroupUnroll = Algo::kGroupUnroll;
        algo.begin_chunk(rows, pp);
#pragma unroll kGroupUnroll
        for (int tt = 0; tt < CHUNK_TICKS; tt += G) {
          uint32_t ped[G], wav[G];
          algo.template group<G, DUMP>(rows + tt * (kWibEthRowBytes / 4), pp, ctx, t0 + tt, ped, wav, tt + G < CHUNK_TICKS);
          if constexpr (DUMP) {
#pragma unroll
            for (int g = 0; g < G; ++g) {
              const size_t o = ((size_t(link) * p.units_stride + unit) * 64 + size_t(t0 + tt + g)) * 32 + lane; // u32 = 2 channels
              if (p.pedestal_out)
                reinterpret_cast<uint32_t*>(p.pedestal_out)[o] = ped[g];
              if (p.waveform_out)
                reinterpret_cast<uint32_t*>(p.waveform_out)[o] = wav[g];
            }
          }
        }
        // Every lane's loads from this stage have completed (their values were consumed above), so after the warp barrier
        // the stage can be handed back to the copy engine: a read-then-async-write hand-off needs no proxy fence.
        __syncwarp();
        produce();
        if (++stg == NSTAGE) {
          stg = 0;
          phase ^= 1u;
        }
      }
    }

    algo.template finish_link<false>(ctx); // drains the policy's software pipeline (deferred hit bookkeeping of the last group)
    Algo::template flush<false>(hits, p.sink, link_base, link, lane); // records carry unit indices of THIS link
    const uint32_t k_end = algo.phase_after((n_units - u0) * 64u);
    algo.store(st, lane, k_end);
    if (lane == 0)
      p.group_flags[link] = kFlagInitialized | (k_end << 8);
    if (p.parts_log2 != 0u) { // publish the state to the warp that runs the link's next slice; the last slice re-arms the counter
#if SWTPG_SLICE_FENCES == 2
      __threadfence();
#endif
      __syncwarp(); // orders the lanes' stores before lane 0's release, which is cumulative
      if (lane == 0) {
        const uint32_t done = part + 1u == (1u << p.parts_log2) ? 0u : slices_before(link, part, u0) + 1u;
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.link_done + link), "r"(done) : "memory");
      }
    }
  }
  // Last warp out re-arms the cursor for the next launch (launches of one handle are stream-ordered: state is carried).
  if (lane == 0) {
    const uint32_t active = min(warps_total, n_items);
    __threadfence();
    if (atomicAdd(p.link_cursor + 1, 1u) == active - 1u) {
      p.link_cursor[0] = 0u;
      p.link_cursor[1] = 0u;
    }
  }
}

// =====================================================================================================================
// WIBEth kernel, CTA form: 4 consumer warps + 1 producer warp work on a QUAD of 4 links in lock-step (the structure of the
// WIB2 kernel below, where one link is four warps wide). A stage of the CTA-wide ring holds one CHUNK_TICKS-tick chunk of
// each of the four links (four bulk copies signalled on ONE `full` barrier, released through ONE `empty` barrier with an
// arrival per consumer), so a consumer's per-chunk bookkeeping is a barrier wait and an arrive — the ring cursor, the
// source addresses and the link hand-out (first quad by block index, every further one claimed from the device-side cursor
// four links at a time, published through a small FIFO) live in the producer warp, once per quad instead of once per link.
// Ragged batches: a quad runs for the longest of its links; a consumer whose link is shorter (or absent) only keeps step.
// =====================================================================================================================
constexpr int kQuad = 4;
constexpr uint32_t kQuadFifo = 4; // quads in flight: the producer is at most NSTAGE chunks ahead, a quad has >= kChunksPerUnit chunks

template<int NSTAGE, int CHUNK_TICKS>
struct WibEthQuadSmem
{
  static constexpr size_t align16(size_t x) { return (x + 15) & ~size_t(15); }
  static constexpr size_t chunk = size_t(kWibEthRowBytes) * CHUNK_TICKS;
  static constexpr size_t bars = align16(size_t(NSTAGE) * kQuad * chunk);
  static constexpr size_t hits = align16(bars + size_t(NSTAGE) * 16);
  static constexpr size_t aux = hits + size_t(kQuad) * HitStage::kCap * 16;
  static constexpr size_t counts = aux + size_t(kQuad) * HitStage::kCap * 4;
  static constexpr size_t fifo = align16(counts + size_t(kQuad) * 4);
  static constexpr size_t total = fifo + kQuadFifo * 32; // per quad: 4 link ids, its length in units, padding
};

template<class Algo, int NSTAGE, int CHUNK_TICKS, bool DUMP>
__global__ void __launch_bounds__((kQuad + 1) * 32)
wibeth_quad_kernel(const KernelParams p)
{
  static_assert(64 % CHUNK_TICKS == 0, "chunk must divide the frame");
  constexpr uint32_t kChunkBytes = kWibEthRowBytes * CHUNK_TICKS;
  constexpr uint32_t kChunksPerUnit = 64 / CHUNK_TICKS;
  static_assert((NSTAGE + kChunksPerUnit - 1) / kChunksPerUnit + 2 <= kQuadFifo, "quad FIFO too short for this ring geometry");
  extern __shared__ __align__(128) uint8_t smem[];
  using L = WibEthQuadSmem<NSTAGE, CHUNK_TICKS>;
  const uint32_t warp = __shfl_sync(0xFFFFFFFFu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31u;
  uint8_t* stages = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::bars);
  uint64_t* empty = full + NSTAGE;
  volatile uint32_t* fifo = reinterpret_cast<volatile uint32_t*>(smem + L::fifo); // [kQuadFifo][8]: links[4], units of the longest
  HitStage hits;
  hits.buf = reinterpret_cast<uint4*>(smem + L::hits) + size_t(warp) * HitStage::kCap;
  hits.aux = reinterpret_cast<uint32_t*>(smem + L::aux) + size_t(warp) * HitStage::kCap;
  hits.cnt = reinterpret_cast<uint32_t*>(smem + L::counts) + warp;

  auto units_of = [&](uint32_t link) -> uint32_t { return p.n_units ? p.n_units[link] : p.units_stride; };
  auto base_of = [&](uint32_t link) -> const uint8_t* { return p.frames + size_t(link) * p.units_stride * SWTPG_WIBETH_FRAME_BYTES; };

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kQuad);
    }
    fence_mbar_init();
  }
  if (lane == 0 && warp < kQuad)
    *hits.cnt = 0u;
  __syncthreads(); // the only CTA-wide barrier: mbarriers visible before anyone waits on them

  if (warp == kQuad) { // ---- producer warp (one lane) ----
    if (lane != 0)
      return;
    uint32_t slot = 0, round = 0, pushed = 0;
    auto wait_slot = [&]() { // all four consumers released the stage's previous contents
      if (round != 0)
        mbar_wait_producer(&empty[slot], (round - 1u) & 1u);
    };
    uint32_t first_link = blockIdx.x * kQuad; // the CTA's first quad; every further one is claimed from the cursor
    for (bool first = true;;) {
      uint32_t nu[kQuad], longest = 0;
      for (;;) { // next quad that has data
        if (!first)
          first_link = gridDim.x * kQuad + atomicAdd(p.link_cursor, uint32_t(kQuad));
        first = false;
        longest = 0;
        if (first_link >= p.n_links)
          break;
#pragma unroll
        for (int c = 0; c < kQuad; ++c) {
          nu[c] = first_link + c < p.n_links ? units_of(first_link + c) : 0u;
          longest = max(longest, nu[c]);
        }
        if (longest != 0)
          break;
      }
      volatile uint32_t* entry = fifo + (pushed & (kQuadFifo - 1u)) * 8u;
      if (first_link >= p.n_links) { // end marker: a stage that completes without data
        wait_slot();
        entry[4] = 0u;
        entry[5] = 1u;
        mbar_arrive(&full[slot]);
        break;
      }
      const uint8_t* src0 = base_of(first_link) + 32;                             // link c of the quad: + c * link_stride
      const size_t link_stride = size_t(p.units_stride) * SWTPG_WIBETH_FRAME_BYTES;
      for (uint32_t unit = 0; unit < longest; ++unit) {
        for (uint32_t ch = 0; ch < kChunksPerUnit; ++ch) {
          wait_slot();
          if (unit == 0 && ch == 0) { // published before the arrive below releases it to the consumers
            entry[0] = first_link;
            entry[4] = longest;
            entry[5] = 0u;
            ++pushed;
          }
          uint32_t bytes = 0;
#pragma unroll
          for (int c = 0; c < kQuad; ++c)
            bytes += unit < nu[c] ? kChunkBytes : 0u;
          mbar_arrive_expect_tx(&full[slot], bytes);
          const uint8_t* src = src0 + size_t(unit) * SWTPG_WIBETH_FRAME_BYTES + ch * kChunkBytes;
#pragma unroll
          for (int c = 0; c < kQuad; ++c)
            if (unit < nu[c])
              bulk_g2s(stages + (slot * kQuad + c) * kChunkBytes, src + c * link_stride, kChunkBytes, &full[slot]);
          if (++slot == NSTAGE) {
            slot = 0;
            ++round;
          }
        }
      }
    }
    __threadfence(); // last CTA out re-arms the cursor for the next launch
    if (atomicAdd(p.link_cursor + 1, 1u) == gridDim.x - 1u) {
      p.link_cursor[0] = 0u;
      p.link_cursor[1] = 0u;
    }
    return;
  }

  // ---- consumer warps: warp c works on link c of every quad ----
  Algo algo;
  algo.configure(p);
  const PairPos pp = pair_pos(lane);
  TickCtx ctx;
  ctx.p = &p;
  ctx.chan0 = 2 * lane;
  ctx.ts = 0;
  ctx.stage = &hits;

  uint32_t stg = 0, phase = 0; // consumer position in the ring and the full barrier's phase
  for (uint32_t popped = 0;; ++popped) {
    mbar_wait(&full[stg], phase); // first chunk of the next quad, or the end marker
    const volatile uint32_t* entry = fifo + (popped & (kQuadFifo - 1u)) * 8u;
    if (entry[5] != 0u)
      break;
    const uint32_t link = entry[0] + warp, longest = entry[4];
    const uint32_t n_units = link < p.n_links ? units_of(link) : 0u;
    const uint8_t* link_base = base_of(link < p.n_links ? link : 0u);
    uint32_t* st = p.state + size_t(link) * kStateWordsPerGroup;
    uint32_t flags = 0;
    bool need_seed = false;
    if (n_units != 0) {
      flags = p.group_flags[link];
      algo.load(st, lane, flags);
      need_seed = !(flags & kFlagInitialized);
    }
    ctx.link = link;
    ctx.link_base = link_base;

    for (uint32_t unit = 0; unit < longest; ++unit) {
      const bool mine = unit < n_units;
      ctx.tick_base = unit * 64u;
      ctx.unit = unit;
      if constexpr (!std::is_same<Algo, PackedSimpleWibEth>::value && !std::is_same<Algo, PackedSimpleWibEthPipe>::value) {
        if (mine) // DAQEthHeader word 1 = timestamp (docs/README.md:81); the packed path reads it when it flushes hits
          ctx.ts = *reinterpret_cast<const unsigned long long*>(link_base + size_t(unit) * SWTPG_WIBETH_FRAME_BYTES + 8);
      }
#pragma unroll 1
      for (int t0 = 0; t0 < 64; t0 += CHUNK_TICKS) {
        if (unit != 0 || t0 != 0)
          mbar_wait(&full[stg], phase);
        if (mine) {
          const uint32_t* rows = reinterpret_cast<const uint32_t*>(stages + (stg * kQuad + warp) * kChunkBytes);
          if (need_seed) {
            algo.seed(Algo::sample(rows, pp));
            need_seed = false;
          }
          constexpr int G = 4;
          static_assert(CHUNK_TICKS % G == 0, "group must divide the chunk");
          constexpr int kGroupUnroll = Algo::kGroupUnroll;
          algo.begin_chunk(rows, pp);
#pragma unroll kGroupUnroll
          for (int tt = 0; tt < CHUNK_TICKS; tt += G) {
            uint32_t ped[G], wav[G];
            algo.template group<G, DUMP>(rows + tt * (kWibEthRowBytes / 4), pp, ctx, t0 + tt, ped, wav, tt + G < CHUNK_TICKS);
            if constexpr (DUMP) {
#pragma unroll
              for (int g = 0; g < G; ++g) {
                const size_t o = ((size_t(link) * p.units_stride + unit) * 64 + size_t(t0 + tt + g)) * 32 + lane; // u32 = 2 channels
                if (p.pedestal_out)
                  reinterpret_cast<uint32_t*>(p.pedestal_out)[o] = ped[g];
                if (p.waveform_out)
                  reinterpret_cast<uint32_t*>(p.waveform_out)[o] = wav[g];
              }
            }
          }
        }
        // Every lane's loads from this stage have completed (their values were consumed above), so after the warp barrier
        // the stage can be handed back to the copy engine: a read-then-async-write hand-off needs no proxy fence.
        __syncwarp();
        if (lane == 0)
          mbar_arrive(&empty[stg]);
        if (++stg == NSTAGE) {
          stg = 0;
          phase ^= 1u;
        }
      }
    }

    if (n_units != 0) {
      algo.template finish_link<false>(ctx);
      Algo::template flush<false>(hits, p.sink, link_base, link, lane); // records carry unit indices of THIS link
      const uint32_t k_end = algo.phase_after(n_units * 64u);
      algo.store(st, lane, k_end);
      if (lane == 0)
        p.group_flags[link] = kFlagInitialized | (k_end << 8);
    }
  }
}

// =====================================================================================================================
// WIB2 kernel: one CTA of 4 consumer warps + 1 producer warp per link (256 channels), persistent over links. A superchunk
// (12 frames x 472 B) is one 5664-byte bulk copy into a CTA-wide ring; consumer warp w works on channels 64w..64w+63 (112
// bytes of every frame's ADC block). full[s]: the copy engine's complete_tx; empty[s]: one arrival per consumer warp. The
// producer warp (one lane; it sleeps on the empty barriers and issues no other instructions) walks the same link/unit
// CTA's links — the first by block index, every further one claimed from the device-side cursor, published to the consumers
// through a small FIFO — and refills a stage as soon as all four consumers have released it — so no consumer ever waits for
// another consumer, only for data (before: warp 0 refilled between its own superchunks and 13 % of all warp time was spent waiting
// on full barriers, profiles/r01_wib2_simple_ncu_full.txt).
// =====================================================================================================================
constexpr int kWib2FrameWords = SWTPG_WIB2_FRAME_BYTES / 4; // 118
constexpr int kWib2Warps = 4;
constexpr uint32_t kWib2Fifo = 8;           // links in flight: the producer is at most NSTAGE superchunks ahead (+ the end marker)
constexpr uint32_t kNoMoreLinks = 0xFFFFFFFFu;

template<int NSTAGE>
struct Wib2Smem
{
  static constexpr size_t align16(size_t x) { return (x + 15) & ~size_t(15); }
  static constexpr size_t bars = align16(size_t(NSTAGE) * SWTPG_WIB2_SUPERCHUNK_BYTES);
  static constexpr size_t hits = align16(bars + size_t(NSTAGE) * 16);
  static constexpr size_t aux = hits + size_t(kWib2Warps) * HitStage::kCap * 16;
  static constexpr size_t counts = aux + size_t(kWib2Warps) * HitStage::kCap * 4;
  static constexpr size_t fifo = align16(counts + size_t(kWib2Warps) * 4);
  static constexpr size_t total = fifo + kWib2Fifo * 4; // link FIFO, producer -> consumers
};

// Algo::kWib2MinCtas caps the registers so that that many CTAs fit an SM (5 = 25 warps): +1 % for SimpleThreshold, +2 % for
// FIR + IQR, -3.5 % for the running sum, which therefore leaves it at 1 (gpurun_out/var30: measured per policy).
template<class Algo, int NSTAGE, bool DUMP>
__global__ void __launch_bounds__((kWib2Warps + 1) * 32, Algo::kWib2MinCtas)
wib2_kernel(const KernelParams p)
{
  constexpr uint32_t kUnit = SWTPG_WIB2_SUPERCHUNK_BYTES;
  extern __shared__ __align__(128) uint8_t smem[];
  using L = Wib2Smem<NSTAGE>;
  const uint32_t warp = __shfl_sync(0xFFFFFFFFu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31u;
  uint8_t* stages = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::bars);
  uint64_t* empty = full + NSTAGE;
  HitStage hits;
  hits.buf = reinterpret_cast<uint4*>(smem + L::hits) + size_t(warp) * HitStage::kCap;
  hits.aux = reinterpret_cast<uint32_t*>(smem + L::aux) + size_t(warp) * HitStage::kCap;
  hits.cnt = reinterpret_cast<uint32_t*>(smem + L::counts) + warp;

  auto units_of = [&](uint32_t link) -> uint32_t { return p.n_units ? p.n_units[link] : p.units_stride; };
  auto base_of = [&](uint32_t link) -> const uint8_t* { return p.frames + size_t(link) * p.units_stride * kUnit; };

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kWib2Warps);
    }
    fence_mbar_init();
  }
  if (lane == 0 && warp < kWib2Warps)
    *hits.cnt = 0u;
  __syncthreads(); // the only CTA-wide barrier: mbarriers visible before anyone waits on them

  volatile uint32_t* fifo = reinterpret_cast<volatile uint32_t*>(smem + L::fifo);
  if (warp == kWib2Warps) { // producer warp: one lane feeds the ring and decides which links this CTA works on
    if (lane != 0)
      return;
    static_assert(NSTAGE + 2 <= int(kWib2Fifo), "link FIFO too short for this ring depth");
    uint32_t slot = 0, round = 0, pushed = 0;
    auto wait_slot = [&]() { // all four consumers released the stage's previous contents
      if (round != 0)
        mbar_wait_producer(&empty[slot], (round - 1u) & 1u);
    };
    uint32_t link = blockIdx.x; // the CTA's first link; every further one is claimed from the cursor (as in wibeth_kernel)
    for (bool first = true;; first = false) {
      uint32_t n_units = 0;
      for (;;) {
        if (!first)
          link = gridDim.x + atomicAdd(p.link_cursor, 1u);
        first = false;
        if (link >= p.n_links)
          break;
        n_units = units_of(link);
        if (n_units != 0)
          break;
      }
      if (link >= p.n_links) { // end marker: a stage that completes without data
        wait_slot();
        fifo[pushed & (kWib2Fifo - 1u)] = kNoMoreLinks;
        mbar_arrive(&full[slot]);
        break;
      }
      const uint8_t* src = base_of(link);
      for (uint32_t unit = 0; unit < n_units; ++unit, src += kUnit) {
        wait_slot();
        if (unit == 0) // published before the arrive below releases it to the consumers
          fifo[pushed++ & (kWib2Fifo - 1u)] = link;
        mbar_arrive_expect_tx(&full[slot], kUnit);
        bulk_g2s(stages + slot * kUnit, src, kUnit, &full[slot]);
        if (++slot == NSTAGE) {
          slot = 0;
          ++round;
        }
      }
    }
    // last CTA out re-arms the cursor for the next launch
    __threadfence();
    if (atomicAdd(p.link_cursor + 1, 1u) == gridDim.x - 1u) {
      p.link_cursor[0] = 0u;
      p.link_cursor[1] = 0u;
    }
    return;
  }

  Algo algo;
  algo.configure(p);
  const PairPos pp = pair_pos(lane);
  TickCtx ctx;
  ctx.p = &p;
  ctx.chan0 = 64 * warp + 2 * lane;
  ctx.ts = 0;
  ctx.stage = &hits;
  const uint32_t row0 = p.wib2_adc_offset / 4 + warp * 28; // word offset of this warp's 112 bytes inside a frame

  uint32_t stg = 0, phase = 0;       // consumer ring position / full-barrier phase
  for (uint32_t popped = 0;; ++popped) {
    mbar_wait(&full[stg], phase);    // first superchunk of the next link, or the end marker
    const uint32_t link = fifo[popped & (kWib2Fifo - 1u)];
    if (link == kNoMoreLinks)
      break;
    const uint32_t n_units = units_of(link);
    const uint8_t* link_base = base_of(link);
    const uint32_t group = link * kWib2Warps + warp;
    uint32_t* st = p.state + size_t(group) * kStateWordsPerGroup;
    const uint32_t flags = p.group_flags[group];
    algo.load(st, lane, flags);
    bool need_seed = !(flags & kFlagInitialized);
    ctx.link = link;
    ctx.link_base = link_base;

    for (uint32_t unit = 0; unit < n_units; ++unit) {
      ctx.tick_base = unit * 12u;
      ctx.unit = unit;
      if (unit != 0)
        mbar_wait(&full[stg], phase);
      const uint32_t* sc = reinterpret_cast<const uint32_t*>(stages + stg * kUnit);
      if constexpr (!std::is_same<Algo, PackedSimpleWib2>::value)
        ctx.ts = uint64_t(sc[1]) | (uint64_t(sc[2]) << 32); // WIB2Frame::get_timestamp, first frame (:350-351)
      const uint32_t* rows = sc + row0;
      if (need_seed) {
        algo.seed(Algo::sample(rows, pp));
        need_seed = false;
      }
      constexpr int G = 4;
      algo.template begin_chunk<kWib2FrameWords>(rows, pp);
#pragma unroll
      for (int tt = 0; tt < 12; tt += G) {
        uint32_t ped[G], wav[G];
        algo.template group<G, DUMP, kWib2FrameWords, true>(rows + tt * kWib2FrameWords, pp, ctx, tt, ped, wav, tt + G < 12);
        if constexpr (DUMP) {
#pragma unroll
          for (int g = 0; g < G; ++g) {
            const size_t o = ((size_t(link) * p.units_stride + unit) * 12 + size_t(tt + g)) * 128 + warp * 32 + lane; // u32 = 2 channels
            if (p.pedestal_out)
              reinterpret_cast<uint32_t*>(p.pedestal_out)[o] = ped[g];
            if (p.waveform_out)
              reinterpret_cast<uint32_t*>(p.waveform_out)[o] = wav[g];
          }
        }
      }
      __syncwarp(); // this warp is done with the stage
      if (lane == 0)
        mbar_arrive(&empty[stg]);
      if (++stg == NSTAGE) {
        stg = 0;
        phase ^= 1u;
      }
    }

    algo.template finish_link<true>(ctx);
    Algo::template flush<true>(hits, p.sink, link_base, link, lane);
    const uint32_t k_end = algo.phase_after(n_units * 12u);
    algo.store(st, lane, k_end);
    if (lane == 0)
      p.group_flags[group] = kFlagInitialized | (k_end << 8);
  }
}

} // namespace swtpg
