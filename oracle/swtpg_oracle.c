/*
 * TEST INFRASTRUCTURE — CPU oracle of the SWTPG hot path (see swtpg_oracle.h and oracle/README.md).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * Every function cites the reference lines it restates (paths relative to the reference repository root).
 * Style: one channel, one tick at a time, int32 intermediates with explicit 16-bit wrap / saturate helpers, so
 * that each AVX2 intrinsic of the reference maps to one visible line here.
 */
#include "swtpg_oracle.h"

#include <math.h>
#include <string.h>

/* ---- 16-bit lane helpers ------------------------------------------------------------------------------------ */
static inline int16_t
wrap16(int32_t x)
{ /* _mm256_add_epi16 / _mm256_sub_epi16 / _mm256_mullo_epi16 lane result */
  return (int16_t)(uint16_t)(uint32_t)x;
}
static inline int16_t
sat16(int32_t x)
{ /* _mm256_adds_epi16 lane result */
  return (int16_t)(x > 32767 ? 32767 : (x < -32768 ? -32768 : x));
}
static inline int16_t
mulhrs16(int16_t a, int16_t b)
{ /* _mm256_mulhrs_epi16 */
  return wrap16((((int32_t)a * (int32_t)b >> 14) + 1) >> 1);
}
static inline int16_t
abs16(int16_t a)
{ /* _mm256_abs_epi16: |-32768| stays -32768 */
  return a < 0 ? wrap16(-(int32_t)a) : a;
}

/* Lane l of register r holds frame channel 16 r + kPerm[l]: wibeth/tpg/FrameExpand.hpp:166-180 parks ADC 15 in
 * lane 8; pinned by unittest/WIBEthFrameExpansion_test.cxx:111,124 and test/apps/wib2_test_bench.cxx:237-254. */
static const int kPerm[16] = { 0, 1, 2, 3, 4, 5, 6, 7, 15, 8, 9, 10, 11, 12, 13, 14 };
static inline int
pos2chan(int pos)
{
  return (pos & ~15) | kPerm[pos & 15];
}

uint16_t
oracle_unpack14(const uint8_t* row, unsigned c)
{ /* channel c at bits [14c, 14c+14) of the little-endian row: SURVEY A.1; fddetdataformats get_adc */
  const unsigned bit = 14u * c, byte = bit >> 3, sh = bit & 7u;
  const uint32_t v = (uint32_t)row[byte] | ((uint32_t)row[byte + 1] << 8) | ((uint32_t)row[byte + 2] << 16);
  return (uint16_t)((v >> sh) & 0x3FFFu);
}

void
oracle_wibeth_expand(const uint8_t* frame, uint16_t* out)
{ /* wibeth/tpg/FrameExpand.hpp:192-246: register r of tick i lands at index i + r*64, 16 lanes each */
  for (int t = 0; t < 64; ++t) {
    const uint8_t* row = frame + 32 + 112 * t;
    for (int pos = 0; pos < 64; ++pos)
      out[(pos >> 4) * 1024 + 16 * t + (pos & 15)] = oracle_unpack14(row, (unsigned)pos2chan(pos));
  }
}

void
oracle_wib2_expand(const uint8_t* sc, int sel, uint32_t adc_offset, uint16_t* out)
{ /* wib2/tpg/FrameExpand.hpp:193-209: block b of frame f lands at index f + b*12 */
  for (int f = 0; f < 12; ++f) {
    const uint8_t* row = sc + 472 * f + adc_offset;
    for (int pos = 0; pos < 128; ++pos)
      out[((pos >> 4) * 12 + f) * 16 + (pos & 15)] = oracle_unpack14(row, (unsigned)(128 * sel + pos2chan(pos)));
  }
}

int
oracle_firwin_int(int n, double cutoff, int multiplier, int16_t* taps)
{ /* src/wib2/tpg/DesignFIR.cpp:20-68: Hamming-windowed sinc, unit DC gain, times multiplier, round() */
  const double pi = 3.14159265358979323846;
  double v[64], sum = 0;
  if (n > 64)
    return -1;
  const int alpha = n / 2;
  for (int m = 0; m < n; ++m) {
    const double w = 0.54 - 0.46 * cos(2.0 * pi * m / (n - 1));
    const double x = cutoff * (m - alpha);
    const double s = (x == 0) ? 1.0 : sin(pi * x) / (pi * x);
    v[m] = w * s;
    sum += v[m];
  }
  for (int m = 0; m < n; ++m)
    taps[m] = (int16_t)round(multiplier * (v[m] / sum));
  return n;
}

/* ---- frugal streaming ---------------------------------------------------------------------------------------- */
/* wibeth/tpg/UtilsAVX2.hpp:24-74 (identical in wib2/tpg/UtilsAVX2.hpp). mask = lane participates. */
static void
frugal_avx2(int16_t* median, int16_t s, int16_t* accum, int16_t acclimit, int mask)
{
  int32_t to_add = (s > *median) ? 1 : ((s == *median) ? 0 : -1); /* :38-47 */
  if (!mask)
    to_add = 0;                                                   /* :50 */
  *accum = wrap16((int32_t)*accum + to_add);                      /* :52 */
  const int is_gt = *accum > acclimit;                            /* :58 */
  /* :59-60 _mm256_sign_epi16(accum, set1(-1*acclimit)) > acclimit */
  const int16_t b = wrap16(-1 * (int32_t)acclimit);
  const int16_t signed_acc = (b < 0) ? wrap16(-(int32_t)*accum) : ((b == 0) ? 0 : *accum);
  const int is_lt = signed_acc > acclimit;
  int32_t step = 0;
  if (is_gt)
    step = 1; /* :63 */
  if (is_lt)
    step = -1; /* :64 (applied second, wins) */
  if (!mask)
    step = 0;                              /* :67 */
  *median = sat16((int32_t)*median + step); /* :69 adds_epi16 */
  if ((is_lt || is_gt) && mask)
    *accum = 0; /* :72-74 */
}

/* wibeth/tpg/ProcessNaive.hpp:21-38 (same in wib2/tpg/ProcessNaive.hpp:20-37): plain int16 ++/--. */
static void
frugal_naive(int16_t* m, int16_t s, int16_t* acc, int16_t acclimit)
{
  if (s > *m)
    *acc = wrap16(*acc + 1);
  if (s < *m)
    *acc = wrap16(*acc - 1);
  if (*acc > acclimit) {
    *m = wrap16(*m + 1);
    *acc = 0;
  }
  if (*acc < -1 * acclimit) {
    *m = wrap16(*m - 1);
    *acc = 0;
  }
}

/* ---- set-up --------------------------------------------------------------------------------------------------- */
void
oracle_link_init(oracle_link* lk, const swtpg_config* cfg, int flavour)
{
  memset(lk, 0, sizeof(*lk)); /* ChanState ctor zeroes everything: wibeth/tpg/ProcessingInfo.hpp:23-40 */
  lk->cfg = *cfg;
  lk->flavour = flavour;
  lk->n_channels = (cfg->format == SWTPG_FORMAT_WIB2) ? 256 : 64;
  if (lk->cfg.wib2_adc_offset == 0)
    lk->cfg.wib2_adc_offset = 20;
  if (lk->cfg.tap_exponent == 0)
    lk->cfg.tap_exponent = 6; /* wibeth/WIBEthFrameProcessor.hpp:69, wib2/WIB2FrameProcessor.hpp:69 */
  int any = 0;
  for (int i = 0; i < 8; ++i)
    any |= lk->cfg.fir_taps[i];
  if (!any) { /* src/wib2/WIB2FrameProcessor.cpp:93-94 */
    oracle_firwin_int(7, 0.1, 1 << lk->cfg.tap_exponent, lk->cfg.fir_taps);
    lk->cfg.fir_taps[7] = 0;
  }
  for (int c = 0; c < lk->n_channels; ++c)
    lk->ch[c].rs_factor = cfg->rs_memory_factor;
}

void
oracle_link_set_memory_factor(oracle_link* lk, const uint16_t* by_channel)
{
  for (int c = 0; c < lk->n_channels; ++c)
    lk->ch[c].rs_factor = by_channel[c];
}

void
oracle_get_state(const oracle_link* lk, swtpg_channel_state* out)
{
  for (int c = 0; c < lk->n_channels; ++c) {
    const oracle_chan* s = &lk->ch[c];
    swtpg_channel_state* o = &out[c];
    memset(o, 0, sizeof(*o));
    o->pedestal = s->median;
    o->accum = s->accum;
    o->quantile25 = s->q25;
    o->quantile75 = s->q75;
    o->accum25 = s->a25;
    o->accum75 = s->a75;
    o->rs = s->rs;
    o->pedestal_rs = s->median_rs;
    o->accum_rs = s->accum_rs;
    o->rs_memory_factor = s->rs_factor;
    o->prev_was_over = s->prev_over;
    o->hit_charge = s->charge;
    o->hit_tover = s->tover;
    o->hit_peak_adc = s->peak_adc;
    o->hit_peak_time = s->peak_time;
    o->initialized = (uint16_t)lk->initialized;
    memcpy(o->prev_samp, s->ring, sizeof(o->prev_samp));
  }
}

/* ---- TP construction ------------------------------------------------------------------------------------------ */
typedef struct tp_sink
{
  swtpg_tp* out;
  size_t cap;
  long n;
  uint32_t link;
} tp_sink;

/* src/wibeth/WIBEthFrameProcessor.cpp:520-545: accepted iff hit_charge != 0 (the `left`/MAGIC tests are what
 * selects the lane in the 16-lane block; here the lane is already known). */
static void
emit_wibeth(tp_sink* k, uint64_t ts, int chan, int t_end, uint16_t charge, uint16_t tover, uint16_t peak_adc, uint16_t peak_time)
{
  if (!charge)
    return;
  if ((size_t)k->n < k->cap) {
    swtpg_tp* tp = &k->out[k->n];
    tp->time_start = ts + (uint64_t)(32 * ((int64_t)t_end - (int64_t)tover)); /* :523 */
    tp->time_peak = tp->time_start + 32u * (uint64_t)peak_time;               /* :524 */
    tp->time_over_threshold = 32u * (uint32_t)tover;                          /* :542 */
    tp->adc_integral = charge;                                                /* :544 */
    tp->adc_peak = peak_adc;                                                  /* :545 */
    tp->channel = (uint16_t)chan;
    tp->link = k->link;
  }
  k->n++;
}

/* src/wib2/WIB2FrameProcessor.cpp:429-455 */
static void
emit_wib2(tp_sink* k, uint64_t ts, int chan, int t_end, uint16_t charge, uint16_t tover)
{
  if (!charge)
    return;
  if ((size_t)k->n < k->cap) {
    swtpg_tp* tp = &k->out[k->n];
    const uint64_t t_begin = ts + (uint64_t)(32 * ((int64_t)t_end - (int64_t)tover)); /* :431-432 */
    const uint64_t t_stop = ts + (uint64_t)(32 * (int64_t)t_end);                     /* :433 */
    tp->time_start = t_begin;
    tp->time_peak = (t_begin + t_stop) / 2;                                 /* :450 */
    tp->time_over_threshold = (uint32_t)((int64_t)tover * 32);              /* :451 */
    tp->adc_integral = charge;                                              /* :453 */
    tp->adc_peak = (uint16_t)(charge / 20);                                 /* :454 */
    tp->channel = (uint16_t)chan;
    tp->link = k->link;
  }
  k->n++;
}

/* ---- per-tick algorithms ------------------------------------------------------------------------------------- */

/* wibeth/tpg/ProcessAVX2.hpp:77-207 for one lane. Returns the pedestal-subtracted sample. */
static int16_t
tick_simple_wibeth(oracle_link* lk, oracle_chan* s, int16_t raw, int chan, int t, uint64_t ts, tp_sink* k)
{
  int16_t x;
  if (lk->flavour == ORACLE_FLAVOUR_NAIVE) { /* wibeth/tpg/ProcessNaive.hpp:83-132 */
    frugal_naive(&s->median, raw, &s->accum, 10); /* :86 hard-coded limit */
    x = wrap16((int32_t)raw - s->median);
    const int is_over = (int32_t)x > (int32_t)lk->cfg.threshold; /* :93 int compare vs uint16 (H5) */
    if (is_over) {
      int32_t tmp = (int32_t)s->charge + x; /* :97-99 saturate at 32767 (H3) */
      if (tmp > 32767)
        tmp = 32767;
      if ((int32_t)x > (int32_t)s->peak_adc) { /* :100-103 */
        s->peak_adc = (uint16_t)x;
        s->peak_time = s->tover;
      }
      s->charge = (uint16_t)(int16_t)tmp;
      s->tover++;
    }
    if (s->prev_over && !is_over) { /* :107-129 */
      emit_wibeth(k, ts, chan, t, s->charge, s->tover, s->peak_adc, s->peak_time);
      s->charge = s->tover = s->peak_adc = s->peak_time = 0;
    }
    s->prev_over = (uint16_t)is_over;
    return x;
  }
  frugal_avx2(&s->median, raw, &s->accum, lk->cfg.frugal_acc_limit, 1);  /* :82 */
  x = wrap16((int32_t)raw - s->median);                                    /* :85 */
  const int is_over = x > (int16_t)lk->cfg.threshold;                      /* :97-98 signed vs (int16)threshold */
  const int left = s->prev_over && !is_over;                               /* :102 */
  s->charge = (uint16_t)wrap16((int32_t)(int16_t)s->charge + (is_over ? x : 0)); /* :114-115; :118 min is a no-op */
  if (x > (int16_t)s->peak_adc) {                                          /* :134-136, not gated by is_over (H6) */
    s->peak_adc = (uint16_t)x;
    s->peak_time = s->tover;
  }
  s->tover = (uint16_t)sat16((int32_t)(int16_t)s->tover + (is_over ? 1 : 0)); /* :139-140 adds_epi16 */
  if (left) {                                                              /* :154-204 */
    emit_wibeth(k, ts, chan, t, s->charge, s->tover, s->peak_adc, s->peak_time);
    s->charge = s->tover = s->peak_adc = s->peak_time = 0;
  }
  s->prev_over = is_over ? 0xFFFFu : 0; /* :207 */
  return x;
}

/* wibeth/tpg/ProcessAbsRSAVX2.hpp:77-170 and ProcessStandardRSAVX2.hpp (differs at :140-144 only). Returns RS-medianRS. */
static int16_t
tick_rs_wibeth(oracle_link* lk, oracle_chan* s, int16_t raw, int chan, int t, uint64_t ts, tp_sink* k, int standard)
{
  const int16_t L = lk->cfg.frugal_acc_limit;
  frugal_avx2(&s->median, raw, &s->accum, L, 1);
  const int16_t x = wrap16((int32_t)raw - s->median);
  const int16_t first = wrap16((int32_t)s->rs * (int16_t)s->rs_factor); /* mullo(RS, R_factor); RS is the CARRIED RS-medianRS */
  int16_t sum;
  if (standard)
    sum = wrap16((int32_t)first + x); /* ProcessStandardRSAVX2.hpp:143 */
  else
    sum = wrap16((int32_t)first + wrap16((int32_t)abs16(x) * (int16_t)lk->cfg.rs_scale_factor)); /* ProcessAbsRSAVX2.hpp:141-142 */
  int16_t rs = mulhrs16(sum, (int16_t)(32768 / 10)); /* UtilsAVX2.hpp:77-81 */
  frugal_avx2(&s->median_rs, rs, &s->accum_rs, L, 1);
  rs = wrap16((int32_t)rs - s->median_rs);
  s->rs = rs;
  const int is_over = rs > (int16_t)lk->cfg.threshold;
  const int left = s->prev_over && !is_over;
  s->charge = (uint16_t)sat16((int32_t)(int16_t)s->charge + (is_over ? x : 0)); /* adds_epi16: saturating here (H3) */
  if (x > (int16_t)s->peak_adc) {
    s->peak_adc = (uint16_t)x;
    s->peak_time = s->tover;
  }
  s->tover = (uint16_t)sat16((int32_t)(int16_t)s->tover + (is_over ? 1 : 0));
  if (left) {
    emit_wibeth(k, ts, chan, t, s->charge, s->tover, s->peak_adc, s->peak_time);
    s->charge = s->tover = s->peak_adc = s->peak_time = 0;
  }
  s->prev_over = is_over ? 0xFFFFu : 0;
  return rs;
}

/* wib2/tpg/ProcessAVX2.hpp:74-179 for one lane. */
static int16_t
tick_simple_wib2(oracle_link* lk, oracle_chan* s, int16_t raw, int chan, int t, uint64_t ts, tp_sink* k)
{
  frugal_avx2(&s->median, raw, &s->accum, 10, 1); /* :79 fixed limit */
  const int16_t x = wrap16((int32_t)raw - s->median);
  const int is_over = x > (int16_t)lk->cfg.threshold;
  const int left = s->prev_over && !is_over;
  const int16_t add = (int16_t)((is_over ? x : 0) >> lk->cfg.tap_exponent);          /* :110-112 srai */
  s->charge = (uint16_t)sat16((int32_t)(int16_t)s->charge + add);
  s->tover = (uint16_t)sat16((int32_t)(int16_t)s->tover + (is_over ? 1 : 0));
  if (left) {
    emit_wib2(k, ts, chan, t, s->charge, s->tover);
    s->charge = s->tover = 0;
  }
  s->prev_over = is_over ? 0xFFFFu : 0;
  return x;
}

/* FIR + IQR: wib2/tpg/ProcessAVX2FIR.hpp:103-283, 16 positions (one AVX2 register) at a time because the threshold
 * is a 64-bit-lane product over groups of 4 adjacent positions (:208). `chan_of_pos[p]` = frame channel
 * of position p; `raw[p]` the sample. filt_out[p] receives the filter output. */
static void
tick_fir_register(oracle_link* lk, const int* chan_of_pos, const int16_t* raw, int t, uint64_t ts, tp_sink* k, unsigned kk,
                  int16_t* filt_out)
{
  const swtpg_config* cfg = &lk->cfg;
  const int16_t multiplier = (int16_t)(1 << cfg->tap_exponent);
  const int16_t adc_max = (int16_t)(INT16_MAX / multiplier);            /* wib2/tpg/ProcessingInfo.hpp:93 */
  const int16_t sigma_max = (int16_t)((1 << 15) / (multiplier * 5));    /* ProcessAVX2FIR.hpp:36 */
  int16_t sigma[16], filt[16];
  if (lk->flavour == ORACLE_FLAVOUR_NAIVE) { /* wib2/tpg/ProcessNaive.hpp:96-150 */
    for (int p = 0; p < 16; ++p) {
      oracle_chan* s = &lk->ch[chan_of_pos[p]];
      int16_t sample = raw[p];
      if (sample < s->median)
        frugal_naive(&s->q25, sample, &s->a25, 10);
      if (sample > s->median)
        frugal_naive(&s->q75, sample, &s->a75, 10);
      frugal_naive(&s->median, sample, &s->accum, 10);
      const int16_t sg = wrap16((int32_t)s->q75 - s->q25);
      sample = wrap16((int32_t)sample - s->median);
      if (sample > adc_max)
        sample = adc_max;
      int16_t f = 0;
      for (unsigned j = 0; j < 8; ++j)
        f = wrap16((int32_t)f + (int32_t)cfg->fir_taps[j] * s->ring[(j + kk) & 7]);
      s->ring[kk & 7] = sample;
      filt_out[p] = f;
      const int is_over = (int32_t)f > 5 * (int32_t)sg * (int32_t)multiplier; /* :123 */
      if (is_over) {
        int32_t tmp = (int32_t)(int16_t)s->charge + (f >> cfg->tap_exponent);
        if (tmp > 32767)
          tmp = 32767;
        s->charge = (uint16_t)(int16_t)tmp;
        s->tover = (uint16_t)wrap16((int32_t)(int16_t)s->tover + 1);
        s->prev_over = 1;
      }
      if (s->prev_over && !is_over) {
        emit_wib2(k, ts, chan_of_pos[p], t, s->charge, s->tover);
        s->charge = s->tover = 0;
        s->prev_over = 0;
      }
    }
    return;
  }
  for (int p = 0; p < 16; ++p) {
    oracle_chan* s = &lk->ch[chan_of_pos[p]];
    const int16_t r = raw[p];
    const int is_gt = r > s->median, is_lt = r < s->median; /* :108-117 masks from the OLD median */
    frugal_avx2(&s->q25, r, &s->a25, 10, is_lt);             /* :119 */
    frugal_avx2(&s->q75, r, &s->a75, 10, is_gt);             /* :121 */
    frugal_avx2(&s->median, r, &s->accum, 10, 1);            /* :125 */
    int16_t x = wrap16((int32_t)r - s->median);              /* :128 */
    int16_t sg = wrap16((int32_t)s->q75 - s->q25);           /* :131 */
    if (sg > sigma_max)
      sg = sigma_max;                                        /* :134 */
    sigma[p] = sg;
    if (x > adc_max)
      x = adc_max;                                           /* :142 */
    int16_t f = 0;
    for (unsigned j = 0; j < 7; ++j)                         /* :176-196 mullo + add, all wrapping */
      f = wrap16((int32_t)f + wrap16((int32_t)cfg->fir_taps[j] * s->ring[(j + kk) & 7]));
    s->ring[kk & 7] = x;                                     /* :199 AFTER the sum: window excludes s(t), s(t-1) */
    filt[p] = f;
    filt_out[p] = f;
  }
  for (int g = 0; g < 4; ++g) { /* :208 `sigma * info.multiplier * info.threshold` on 4 x int64 lanes */
    uint64_t v = 0;
    for (int j = 0; j < 4; ++j)
      v |= (uint64_t)(uint16_t)sigma[4 * g + j] << (16 * j);
    v = v * (uint64_t)(int64_t)multiplier;
    v = v * (uint64_t)cfg->threshold;
    for (int j = 0; j < 4; ++j) {
      const int p = 4 * g + j;
      oracle_chan* s = &lk->ch[chan_of_pos[p]];
      const int16_t thr = (int16_t)(uint16_t)(v >> (16 * j));
      const int is_over = filt[p] > thr;
      const int left = s->prev_over && !is_over;                                          /* :210 */
      const int16_t add = (int16_t)((is_over ? filt[p] : 0) >> cfg->tap_exponent);        /* :221-223 */
      s->charge = (uint16_t)sat16((int32_t)(int16_t)s->charge + add);
      s->tover = (uint16_t)sat16((int32_t)(int16_t)s->tover + (is_over ? 1 : 0));         /* :236-237 */
      if (left) {                                                                         /* :251-281 */
        emit_wib2(k, ts, chan_of_pos[p], t, s->charge, s->tover);
        s->charge = s->tover = 0;
      }
      s->prev_over = is_over ? 0xFFFFu : 0;
    }
  }
}

/* WIB2 AbsRS: wib2/tpg/ProcessRSAVX2.hpp:24-330, 16 positions (one AVX2 register) at a time because the threshold
 * `sigma * info.threshold` (:186) is again a product on 4 x int64 lanes. R = 8 and scale = 5 are literals (:29-33); all
 * four frugal trackers use L = 10; sigmaMax = 2^15 / (multiplier * threshold) (:36); the charge accumulates
 * adds(RS, medianRS) >> tap_exponent (:200-203). level_out[p] receives RS - medianRS. */
static void
tick_absrs_wib2_register(oracle_link* lk, const int* chan_of_pos, const int16_t* raw, int t, uint64_t ts, tp_sink* k, int16_t* level_out)
{
  const swtpg_config* cfg = &lk->cfg;
  const int multiplier = 1 << cfg->tap_exponent;
  const int16_t sigma_max = (int16_t)((1 << 15) / (multiplier * (int)cfg->threshold)); /* threshold >= 1 (else the reference divides by 0) */
  int16_t sigma[16], rsv[16];
  for (int p = 0; p < 16; ++p) {
    oracle_chan* s = &lk->ch[chan_of_pos[p]];
    const int16_t r = raw[p];
    const int is_gt = r > s->median, is_lt = r < s->median;   /* :116-124 masks from the OLD median */
    frugal_avx2(&s->q25, r, &s->a25, 10, is_lt);               /* :126 */
    frugal_avx2(&s->q75, r, &s->a75, 10, is_gt);               /* :128 */
    frugal_avx2(&s->median, r, &s->accum, 10, 1);              /* :132 */
    const int16_t x = wrap16((int32_t)r - s->median);          /* :135 */
    const int16_t first = wrap16((int32_t)s->rs * 8);          /* :151 mullo(RS, R_factor) */
    const int16_t second = wrap16((int32_t)abs16(x) * 5);      /* :154 */
    int16_t rs = mulhrs16(wrap16((int32_t)first + second), (int16_t)(32768 / 10)); /* :158, UtilsAVX2.hpp:77-81 */
    frugal_avx2(&s->median_rs, rs, &s->accum_rs, 10, 1);       /* :169 */
    rs = wrap16((int32_t)rs - s->median_rs);                   /* :173 */
    s->rs = rs;
    rsv[p] = rs;
    level_out[p] = rs;
    int16_t sg = wrap16((int32_t)s->q75 - s->q25);             /* :184 */
    if (sg > sigma_max)
      sg = sigma_max;                                          /* :188 */
    sigma[p] = sg;
  }
  for (int g = 0; g < 4; ++g) { /* :198 `sigma * info.threshold` on 4 x int64 lanes */
    uint64_t v = 0;
    for (int j = 0; j < 4; ++j)
      v |= (uint64_t)(uint16_t)sigma[4 * g + j] << (16 * j);
    v = v * (uint64_t)cfg->threshold;
    for (int j = 0; j < 4; ++j) {
      const int p = 4 * g + j;
      oracle_chan* s = &lk->ch[chan_of_pos[p]];
      const int16_t thr = (int16_t)(uint16_t)(v >> (16 * j));
      const int is_over = rsv[p] > thr;
      const int left = s->prev_over && !is_over;                                           /* :200 */
      const int16_t temp = sat16((int32_t)rsv[p] + s->median_rs);                          /* :210 adds_epi16(RS, medianRS) */
      const int16_t add = (int16_t)((is_over ? temp : 0) >> cfg->tap_exponent);            /* :211-213 */
      s->charge = (uint16_t)sat16((int32_t)(int16_t)s->charge + add);
      s->tover = (uint16_t)sat16((int32_t)(int16_t)s->tover + (is_over ? 1 : 0));          /* :230-231 */
      if (left) {
        emit_wib2(k, ts, chan_of_pos[p], t, s->charge, s->tover);
        s->charge = s->tover = 0;
      }
      s->prev_over = is_over ? 0xFFFFu : 0;
    }
  }
}

/* ---- driver ---------------------------------------------------------------------------------------------------- */
long
oracle_process(oracle_link* lk, const uint8_t* units, size_t n_units, uint32_t link_id, swtpg_tp* out, size_t cap,
               int16_t* pedestal_out, int16_t* waveform_out)
{
  const swtpg_config* cfg = &lk->cfg;
  const int wib2 = cfg->format == SWTPG_FORMAT_WIB2;
  const int nch = lk->n_channels;
  const int nticks = wib2 ? 12 : 64;
  const size_t unit_bytes = wib2 ? SWTPG_WIB2_SUPERCHUNK_BYTES : SWTPG_WIBETH_FRAME_BYTES;
  tp_sink k = { out, cap, 0, link_id };

  for (size_t u = 0; u < n_units; ++u) {
    const uint8_t* unit = units + u * unit_bytes;
    uint64_t ts;
    if (wib2) { /* WIB2Frame::get_timestamp of the first frame: src/wib2/WIB2FrameProcessor.cpp:350-351 */
      uint32_t w[2];
      memcpy(w, unit + 4, 8);
      ts = (uint64_t)w[0] | ((uint64_t)w[1] << 32);
    } else {    /* src/wibeth/WIBEthFrameProcessor.cpp:415-416; 2nd 64-bit word (docs/README.md:81) */
      memcpy(&ts, unit + 8, 8);
    }
    for (int t = 0; t < nticks; ++t) {
      const uint8_t* row = wib2 ? unit + 472 * t + cfg->wib2_adc_offset : unit + 32 + 112 * t;
      if (!lk->initialized) { /* setState: pedestal = first sample; quartiles +-20 (wibeth/tpg/ProcessingInfo.hpp:116-144) */
        for (int c = 0; c < nch; ++c) {
          const int16_t ped = (int16_t)oracle_unpack14(row, (unsigned)c);
          lk->ch[c].median = ped;
          lk->ch[c].q25 = wrap16(ped - 20);
          lk->ch[c].q75 = wrap16(ped + 20);
        }
        lk->initialized = 1;
      }
      int16_t* ped_row = pedestal_out ? pedestal_out + ((u * nticks + t) * (size_t)nch) : 0;
      int16_t* wav_row = waveform_out ? waveform_out + ((u * nticks + t) * (size_t)nch) : 0;
      if (wib2 && cfg->algorithm == SWTPG_ALGO_ABS_RS) {
        for (int r = 0; r < nch / 16; ++r) {
          int chan_of_pos[16];
          int16_t raw[16], lvl[16];
          for (int p = 0; p < 16; ++p) {
            chan_of_pos[p] = 16 * r + kPerm[p];
            raw[p] = (int16_t)oracle_unpack14(row, (unsigned)chan_of_pos[p]);
          }
          tick_absrs_wib2_register(lk, chan_of_pos, raw, t, ts, &k, lvl);
          for (int p = 0; p < 16; ++p) {
            if (ped_row)
              ped_row[chan_of_pos[p]] = lk->ch[chan_of_pos[p]].median;
            if (wav_row)
              wav_row[chan_of_pos[p]] = lvl[p];
          }
        }
      } else if (cfg->algorithm == SWTPG_ALGO_FIR_IQR) {
        const unsigned kk = (lk->k0 + (unsigned)t) & 7u;
        for (int r = 0; r < nch / 16; ++r) {
          int chan_of_pos[16];
          int16_t raw[16], filt[16];
          for (int p = 0; p < 16; ++p) {
            chan_of_pos[p] = 16 * r + kPerm[p];
            raw[p] = (int16_t)oracle_unpack14(row, (unsigned)chan_of_pos[p]);
          }
          tick_fir_register(lk, chan_of_pos, raw, t, ts, &k, kk, filt);
          for (int p = 0; p < 16; ++p) {
            if (ped_row)
              ped_row[chan_of_pos[p]] = lk->ch[chan_of_pos[p]].median;
            if (wav_row)
              wav_row[chan_of_pos[p]] = filt[p];
          }
        }
      } else {
        for (int c = 0; c < nch; ++c) {
          const int16_t raw = (int16_t)oracle_unpack14(row, (unsigned)c);
          oracle_chan* s = &lk->ch[c];
          int16_t w;
          if (cfg->algorithm == SWTPG_ALGO_SIMPLE_THRESHOLD)
            w = wib2 ? tick_simple_wib2(lk, s, raw, c, t, ts, &k) : tick_simple_wibeth(lk, s, raw, c, t, ts, &k);
          else
            w = tick_rs_wibeth(lk, s, raw, c, t, ts, &k, cfg->algorithm == SWTPG_ALGO_STANDARD_RS);
          if (ped_row)
            ped_row[c] = s->median;
          if (wav_row)
            wav_row[c] = w;
        }
      }
    }
    lk->k0 = (lk->k0 + (unsigned)nticks) & 7u; /* wib2/tpg/ProcessAVX2FIR.hpp:306 */
  }
  return k.n;
}
