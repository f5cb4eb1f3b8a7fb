/*
 * swtpg_framegen.h — synthetic WIBEth / WIB2 frame generator exported by libswtpg_framegen.so (its own small library).
 *
 * Test / benchmark utility, NOT a reference interface (the reference replays recorded files through an emulator that
 * is absent from the snapshot, docs/README.md:20-48). Host and device variants produce byte-identical frames
 * (fdreadoutlibs_b200/csrc/framegen.h is compiled into both), so the CPU checkers and the GPU see the same bytes.
 *
 * Output layout: link-major, unit u of link l at ((l * n_units) + u) * unit_bytes, the layout swtpg_process_* take.
 * Link l covers global channels (link0 + l) * C .. + C - 1; unit u covers absolute ticks (unit0 + u) * T .. + T - 1
 * and carries timestamp ts0 + (unit0 + u) * T * 32 (WIBEth: C=64, T=64; WIB2: C=256, T=12, 12 frames 32 apart).
 */
#ifndef SWTPG_FRAMEGEN_API_H_
#define SWTPG_FRAMEGEN_API_H_

#include "swtpg.h"
#include "../fdreadoutlibs_b200/csrc/framegen.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Noise sigma 5 ADC, pedestal 900 + 7*(gch % 97), pulses U[40,400] ADC x half width U[3,10], bipolar on 2/3 of the
 * channels; pulse_prob_q32 = pulses_per_64_ticks * 2^32. */
void swtpg_gen_default_params(swtpg_gen_params* p, uint64_t seed, double pulses_per_64_ticks);

swtpg_status swtpg_gen_wibeth_host(const swtpg_gen_params* p, uint32_t link0, uint32_t n_links, uint64_t unit0,
                                   uint32_t n_units, uint64_t ts0, void* out, int n_threads);
swtpg_status swtpg_gen_wib2_host(const swtpg_gen_params* p, uint32_t link0, uint32_t n_links, uint64_t unit0, uint32_t n_units,
                                 uint64_t ts0, uint32_t adc_offset, void* out, int n_threads);
/* d_out is device memory on the current device; `stream` a cudaStream_t or NULL. Asynchronous. */
swtpg_status swtpg_gen_wibeth_device(const swtpg_gen_params* p, uint32_t link0, uint32_t n_links, uint64_t unit0,
                                     uint32_t n_units, uint64_t ts0, void* d_out, void* stream);
swtpg_status swtpg_gen_wib2_device(const swtpg_gen_params* p, uint32_t link0, uint32_t n_links, uint64_t unit0,
                                   uint32_t n_units, uint64_t ts0, uint32_t adc_offset, void* d_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif
