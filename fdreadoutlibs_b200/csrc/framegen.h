/*
 * framegen.h — synthetic TPC waveform / frame generator shared by host (gcc, g++) and device (nvcc) code.
 *
 * Test and benchmark utility, not a reference interface: the reference replays recorded frame files through an
 * emulator that is not in the snapshot (docs/README.md:20-48). Everything is integer arithmetic on a counter-based
 * hash keyed by (seed, global channel, absolute tick), so any sharding of links over threads, GPUs or ranks yields
 * byte-identical frames, and the CPU checkers see exactly what the GPU saw.
 *
 * Waveform model: per-channel pedestal ped_base + ped_step*(gch % ped_mod); noise ~ N(0, sigma)
 * approximated by a centred sum of 8 uniform bytes (Irwin-Hall, sd 209.02) scaled to sigma = noise_q8/256 ADC;
 * pulses start in a 64-tick block with probability pulse_prob_q32 / 2^32, triangular, amplitude U[amp_min, amp_max],
 * half-width U[hw_min, hw_max] ticks; on channels with (gch % 3) != 0 (induction-like) the triangle is followed by
 * a negative lobe of half the amplitude. Samples are clipped to [0, 16383].
 */
#ifndef SWTPG_FRAMEGEN_H_
#define SWTPG_FRAMEGEN_H_

#include <stdint.h>

#ifdef __CUDACC__
#define SWTPG_HD __host__ __device__ __forceinline__
#else
#define SWTPG_HD static inline
#endif

typedef struct swtpg_gen_params
{
  uint64_t seed;
  uint32_t noise_q8;       /* noise sigma in 1/256 ADC (5.0 ADC -> 1280) */
  uint32_t pulse_prob_q32; /* P(pulse starts in a given 64-tick block of a channel) * 2^32 */
  uint16_t amp_min, amp_max;
  uint16_t hw_min, hw_max; /* half width in ticks; 4*hw_max <= 64 */
  uint16_t ped_base, ped_step, ped_mod;
  uint16_t bipolar; /* 1: channels with gch % 3 != 0 get a negative second lobe */
} swtpg_gen_params;

SWTPG_HD uint64_t
swtpg_mix64(uint64_t x)
{ /* splitmix64 finaliser */
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

SWTPG_HD int32_t
swtpg_gen_noise(const swtpg_gen_params* p, uint64_t gch, uint64_t tick)
{
  uint64_t h = swtpg_mix64(p->seed ^ swtpg_mix64(gch * 0x100000001B3ull + tick));
  /* sum of the 8 bytes of h: SWAR */
  uint64_t s = (h & 0x00FF00FF00FF00FFull) + ((h >> 8) & 0x00FF00FF00FF00FFull);
  s = (s & 0x0000FFFF0000FFFFull) + ((s >> 16) & 0x0000FFFF0000FFFFull);
  int64_t c = (int64_t)((s & 0xFFFFFFFFull) + (s >> 32)) - 1020;
  /* c has sd 209.02; 2^26 / (209.02 * 256) = 1254.1 */
  return (int32_t)((c * (int64_t)p->noise_q8 * 1254 + (1 << 25)) >> 26);
}

/* Contribution at absolute tick `tick` of the pulse (if any) that starts in 64-tick block `blk` of channel gch. */
SWTPG_HD int32_t
swtpg_gen_pulse(const swtpg_gen_params* p, uint64_t gch, uint64_t blk, uint64_t tick)
{
  uint64_t h = swtpg_mix64((p->seed * 0x9E3779B97F4A7C15ull) ^ swtpg_mix64(gch * 0x1000193ull + blk * 0xD6E8FEB86659FD93ull));
  if ((uint32_t)h >= p->pulse_prob_q32)
    return 0;
  int64_t start = (int64_t)(blk * 64 + ((h >> 32) & 63));
  int32_t amp = (int32_t)p->amp_min + (int32_t)((h >> 38) % (uint64_t)(p->amp_max - p->amp_min + 1));
  int32_t hw = (int32_t)p->hw_min + (int32_t)((h >> 52) % (uint64_t)(p->hw_max - p->hw_min + 1));
  int64_t d = (int64_t)tick - start;
  if (d < 0 || d > 4 * hw)
    return 0;
  if (d <= 2 * hw) {
    int32_t a = (int32_t)d - hw;
    if (a < 0)
      a = -a;
    return amp * (hw - a) / hw;
  }
  if (!p->bipolar || (gch % 3) == 0)
    return 0;
  int32_t a = (int32_t)d - 3 * hw;
  if (a < 0)
    a = -a;
  return -((amp * (hw - a) / hw) / 2);
}

SWTPG_HD uint16_t
swtpg_gen_sample(const swtpg_gen_params* p, uint64_t gch, uint64_t tick)
{
  int32_t v = (int32_t)p->ped_base + (int32_t)p->ped_step * (int32_t)(gch % (uint64_t)p->ped_mod);
  v += swtpg_gen_noise(p, gch, tick);
  uint64_t blk = tick >> 6;
  v += swtpg_gen_pulse(p, gch, blk, tick);
  if (blk > 0)
    v += swtpg_gen_pulse(p, gch, blk - 1, tick);
  if (v < 0)
    v = 0;
  if (v > 16383)
    v = 16383;
  return (uint16_t)v;
}

/* Word 0 of the WIBEth DAQEthHeader (bit positions restated from fddetdataformats, unpinned by the reference). */
SWTPG_HD uint64_t
swtpg_wibeth_header_word0(uint32_t det_id, uint32_t crate, uint32_t slot, uint32_t stream, uint32_t seq)
{
  return (uint64_t)2 | ((uint64_t)(det_id & 0x3F) << 6) | ((uint64_t)(crate & 0x3FF) << 12) | ((uint64_t)(slot & 0xF) << 22) |
         ((uint64_t)(stream & 0xFF) << 26) | ((uint64_t)(seq & 0xFFF) << 40) | ((uint64_t)0x382 << 52);
}

/* One 112-byte tick row (14 u64 words) of a WIBEth frame: channels 64*link .. 64*link+63 at absolute tick `tick`. */
SWTPG_HD void
swtpg_gen_wibeth_row(const swtpg_gen_params* p, uint64_t link, uint64_t tick, uint64_t* row /* 14 words */)
{
  uint64_t acc = 0;
  int nbits = 0, w = 0;
  for (int c = 0; c < 64; ++c) {
    uint64_t v = swtpg_gen_sample(p, link * 64 + (uint64_t)c, tick);
    acc |= v << nbits;
    nbits += 14;
    if (nbits >= 64) {
      row[w++] = acc;
      nbits -= 64;
      acc = nbits ? (v >> (14 - nbits)) : 0;
    }
  }
}

/* One 448-byte ADC block (112 u32 words) of a WIB2 frame: channels 256*link .. +255 at absolute tick `tick`. */
SWTPG_HD void
swtpg_gen_wib2_adcs(const swtpg_gen_params* p, uint64_t link, uint64_t tick, uint32_t* words /* 112 */)
{
  uint64_t acc = 0;
  int nbits = 0, w = 0;
  for (int c = 0; c < 256; ++c) {
    uint64_t v = swtpg_gen_sample(p, link * 256 + (uint64_t)c, tick);
    acc |= v << nbits;
    nbits += 14;
    if (nbits >= 32) {
      words[w++] = (uint32_t)acc;
      acc >>= 32;
      nbits -= 32;
    }
  }
}

#endif /* SWTPG_FRAMEGEN_H_ */
