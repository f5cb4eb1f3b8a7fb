// Synthetic frame generator entry points (include/swtpg_framegen.h). Test / benchmark utility.
#include "../../include/swtpg_framegen.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

namespace {

constexpr uint32_t kDetId = 3, kCrate = 1;

inline void
wibeth_fill_unit(const swtpg_gen_params& p, uint64_t link, uint64_t unit, uint64_t ts0, uint8_t* dst)
{
  uint64_t hdr[4] = { swtpg_wibeth_header_word0(kDetId, kCrate, uint32_t(link / 8) & 0xF, uint32_t(link & 0xFF), uint32_t(unit)),
                      ts0 + unit * 2048ull, 0, 0 };
  memcpy(dst, hdr, 32);
  for (uint64_t t = 0; t < 64; ++t) {
    uint64_t row[14];
    swtpg_gen_wibeth_row(&p, link, unit * 64 + t, row);
    memcpy(dst + 32 + 112 * t, row, 112);
  }
}

inline void
wib2_fill_unit(const swtpg_gen_params& p, uint64_t link, uint64_t unit, uint64_t ts0, uint32_t adc_offset, uint8_t* dst)
{
  for (uint64_t f = 0; f < 12; ++f) {
    uint8_t* fr = dst + 472 * f;
    memset(fr, 0, 472);
    const uint64_t ts = ts0 + (unit * 12 + f) * 32ull;
    uint32_t hdr[3] = { 4u | (kDetId << 6) | (kCrate << 12) | (uint32_t(link & 0x3F) << 26), uint32_t(ts), uint32_t(ts >> 32) };
    memcpy(fr, hdr, 12);
    uint32_t words[112];
    swtpg_gen_wib2_adcs(&p, link, unit * 12 + f, words);
    memcpy(fr + adc_offset, words, 448);
  }
}

template<typename F>
void
parallel_for(size_t n, int n_threads, F f)
{
  n_threads = std::max(1, std::min<int>(n_threads, int(n)));
  if (n_threads == 1) {
    for (size_t i = 0; i < n; ++i)
      f(i);
    return;
  }
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; ++t)
    th.emplace_back([=]() {
      for (size_t i = size_t(t); i < n; i += size_t(n_threads))
        f(i);
    });
  for (auto& x : th)
    x.join();
}

// one thread per (link, unit, tick): 14 u64 words of a row (+ header written by tick 0)
__global__ void
gen_wibeth_kernel(swtpg_gen_params p, uint32_t link0, uint32_t n_links, uint64_t unit0, uint32_t n_units, uint64_t ts0, uint8_t* out)
{
  const size_t gid = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = size_t(n_links) * n_units * 64;
  if (gid >= total)
    return;
  const uint32_t t = uint32_t(gid & 63);
  const size_t lu = gid >> 6;
  const uint32_t u = uint32_t(lu % n_units);
  const uint32_t l = uint32_t(lu / n_units);
  const uint64_t link = uint64_t(link0) + l, unit = unit0 + u;
  uint8_t* dst = out + lu * 7200;
  uint64_t row[14];
  swtpg_gen_wibeth_row(&p, link, unit * 64 + t, row);
  uint64_t* d64 = reinterpret_cast<uint64_t*>(dst + 32 + 112 * t);
#pragma unroll
  for (int i = 0; i < 14; ++i)
    d64[i] = row[i];
  if (t == 0) {
    uint64_t* h = reinterpret_cast<uint64_t*>(dst);
    h[0] = swtpg_wibeth_header_word0(kDetId, kCrate, uint32_t(link / 8) & 0xF, uint32_t(link & 0xFF), uint32_t(unit));
    h[1] = ts0 + unit * 2048ull;
    h[2] = 0;
    h[3] = 0;
  }
}

// one thread per (link, unit, frame): 112 u32 words (+ header/trailer)
__global__ void
gen_wib2_kernel(swtpg_gen_params p, uint32_t link0, uint32_t n_links, uint64_t unit0, uint32_t n_units, uint64_t ts0,
                uint32_t adc_offset, uint8_t* out)
{
  const size_t gid = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = size_t(n_links) * n_units * 12;
  if (gid >= total)
    return;
  const uint32_t f = uint32_t(gid % 12);
  const size_t lu = gid / 12;
  const uint32_t u = uint32_t(lu % n_units);
  const uint32_t l = uint32_t(lu / n_units);
  const uint64_t link = uint64_t(link0) + l, unit = unit0 + u;
  uint32_t* fr = reinterpret_cast<uint32_t*>(out + lu * 5664 + 472 * f);
  for (int i = 0; i < 118; ++i)
    fr[i] = 0;
  const uint64_t ts = ts0 + (unit * 12 + f) * 32ull;
  fr[0] = 4u | (kDetId << 6) | (kCrate << 12) | (uint32_t(link & 0x3F) << 26);
  fr[1] = uint32_t(ts);
  fr[2] = uint32_t(ts >> 32);
  uint32_t words[112];
  swtpg_gen_wib2_adcs(&p, link, unit * 12 + f, words);
  uint32_t* adc = fr + adc_offset / 4;
  for (int i = 0; i < 112; ++i)
    adc[i] = words[i];
}

} // namespace

extern "C" {

void
swtpg_gen_default_params(swtpg_gen_params* p, uint64_t seed, double pulses_per_64_ticks)
{
  memset(p, 0, sizeof(*p));
  p->seed = seed;
  p->noise_q8 = 5 * 256;
  double q = pulses_per_64_ticks * 4294967296.0;
  p->pulse_prob_q32 = q >= 4294967295.0 ? 0xFFFFFFFFu : (q <= 0 ? 0u : uint32_t(q));
  p->amp_min = 40;
  p->amp_max = 400;
  p->hw_min = 3;
  p->hw_max = 10;
  p->ped_base = 900;
  p->ped_step = 7;
  p->ped_mod = 97;
  p->bipolar = 1;
}

swtpg_status
swtpg_gen_wibeth_host(const swtpg_gen_params* p, uint32_t link0, uint32_t n_links, uint64_t unit0, uint32_t n_units, uint64_t ts0,
                      void* out, int n_threads)
{
  if (!p || !out || p->ped_mod == 0 || p->amp_max < p->amp_min || p->hw_max < p->hw_min || p->hw_min == 0 || 4 * p->hw_max > 64)
    return SWTPG_ERR_INVALID_ARG;
  const swtpg_gen_params pp = *p;
  uint8_t* o = static_cast<uint8_t*>(out);
  parallel_for(size_t(n_links) * n_units, n_threads, [=](size_t i) {
    wibeth_fill_unit(pp, uint64_t(link0) + i / n_units, unit0 + i % n_units, ts0, o + i * 7200);
  });
  return SWTPG_OK;
}

swtpg_status
swtpg_gen_wib2_host(const swtpg_gen_params* p, uint32_t link0, uint32_t n_links, uint64_t unit0, uint32_t n_units, uint64_t ts0,
                    uint32_t adc_offset, void* out, int n_threads)
{
  if (adc_offset == 0)
    adc_offset = 20;
  if (!p || !out || p->ped_mod == 0 || p->amp_max < p->amp_min || p->hw_max < p->hw_min || p->hw_min == 0 || 4 * p->hw_max > 64 ||
      adc_offset % 4 || adc_offset < 12 || adc_offset + 448 > 472)
    return SWTPG_ERR_INVALID_ARG;
  const swtpg_gen_params pp = *p;
  uint8_t* o = static_cast<uint8_t*>(out);
  parallel_for(size_t(n_links) * n_units, n_threads, [=](size_t i) {
    wib2_fill_unit(pp, uint64_t(link0) + i / n_units, unit0 + i % n_units, ts0, adc_offset, o + i * 5664);
  });
  return SWTPG_OK;
}

swtpg_status
swtpg_gen_wibeth_device(const swtpg_gen_params* p, uint32_t link0, uint32_t n_links, uint64_t unit0, uint32_t n_units,
                        uint64_t ts0, void* d_out, void* stream)
{
  if (!p || !d_out || p->ped_mod == 0 || p->amp_max < p->amp_min || p->hw_max < p->hw_min || p->hw_min == 0 || 4 * p->hw_max > 64)
    return SWTPG_ERR_INVALID_ARG;
  const size_t total = size_t(n_links) * n_units * 64;
  if (total == 0)
    return SWTPG_OK;
  gen_wibeth_kernel<<<unsigned((total + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
    *p, link0, n_links, unit0, n_units, ts0, static_cast<uint8_t*>(d_out));
  return cudaGetLastError() == cudaSuccess ? SWTPG_OK : SWTPG_ERR_CUDA;
}

swtpg_status
swtpg_gen_wib2_device(const swtpg_gen_params* p, uint32_t link0, uint32_t n_links, uint64_t unit0, uint32_t n_units, uint64_t ts0,
                      uint32_t adc_offset, void* d_out, void* stream)
{
  if (adc_offset == 0)
    adc_offset = 20;
  if (!p || !d_out || p->ped_mod == 0 || p->amp_max < p->amp_min || p->hw_max < p->hw_min || p->hw_min == 0 || 4 * p->hw_max > 64 ||
      adc_offset % 4 || adc_offset < 12 || adc_offset + 448 > 472)
    return SWTPG_ERR_INVALID_ARG;
  const size_t total = size_t(n_links) * n_units * 12;
  if (total == 0)
    return SWTPG_OK;
  gen_wib2_kernel<<<unsigned((total + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
    *p, link0, n_links, unit0, n_units, ts0, adc_offset, static_cast<uint8_t*>(d_out));
  return cudaGetLastError() == cudaSuccess ? SWTPG_OK : SWTPG_ERR_CUDA;
}

} // extern "C"
