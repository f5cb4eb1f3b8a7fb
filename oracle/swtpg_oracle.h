/*
 * TEST INFRASTRUCTURE — CPU oracle of the SWTPG hot path. Not part of the product; see oracle/README.md.
 *
 * Plain-C, one-channel-at-a-time restatement of the reference's AVX2 arithmetic (and, behind `flavour`, of the
 * places where its scalar "naive" code differs). PINNED: tests/test_oracle_vs_reference.py checks it against the
 * reference's own headers compiled here (oracle/_ref/libswtpg_ref.so), against the unpack permutation test
 * (unittest/WIBEthFrameExpansion_test.cxx:92-156) and against the golden TPs of docs/README.md:86-88,136-146;
 * tests/golden/ holds reference-generated vectors so the same check runs where /root/reference is absent.
 */
#ifndef SWTPG_ORACLE_H_
#define SWTPG_ORACLE_H_

#include "../include/swtpg.h"

#ifdef __cplusplus
extern "C" {
#endif

enum
{
  ORACLE_FLAVOUR_AVX2 = 0, /* canonical: what production runs (src/wibeth/WIBEthFrameProcessor.cpp:184) */
  ORACLE_FLAVOUR_NAIVE = 1 /* ProcessNaive.hpp semantics where they differ */
};

typedef struct oracle_chan
{
  int16_t median, accum;
  int16_t q25, q75, a25, a75;
  int16_t rs, median_rs, accum_rs;
  uint16_t rs_factor;
  uint16_t prev_over, charge, tover, peak_adc, peak_time;
  int16_t ring[8];
} oracle_chan;

typedef struct oracle_link
{
  swtpg_config cfg;
  int flavour;
  int n_channels;
  int initialized; /* first unit seen (setState done) */
  unsigned k0;     /* absTimeModNTAPS */
  oracle_chan ch[SWTPG_WIB2_CHANNELS];
} oracle_link;

/* Zeroed state, as FrameProcessor::start leaves it. cfg->n_links / max_units / device are ignored. */
void oracle_link_init(oracle_link* lk, const swtpg_config* cfg, int flavour);
/* Per-channel RS memory factor (by frame channel); default cfg->rs_memory_factor. */
void oracle_link_set_memory_factor(oracle_link* lk, const uint16_t* by_channel);

/* Run n_units consecutive units (7200-B frames or 5664-B superchunks) of one link. TPs appended to out[0..cap);
 * returns the number found (may exceed cap: the excess is counted, not stored). pedestal_out / waveform_out
 * (optional) receive int16 [unit][tick][channel]: the pedestal after its update, and the waveform the threshold
 * is applied to. */
long oracle_process(oracle_link* lk, const uint8_t* units, size_t n_units, uint32_t link_id, swtpg_tp* out, size_t cap,
                    int16_t* pedestal_out, int16_t* waveform_out);

/* 14-bit field c of a little-endian packed row. */
uint16_t oracle_unpack14(const uint8_t* row, unsigned c);
/* The reference's expanded register layout of one WIBEth frame: out[r*1024 + 16 t + lane]
 * (wibeth/tpg/FrameExpand.hpp:192-246), for the unpack known-answer test. */
void oracle_wibeth_expand(const uint8_t* frame, uint16_t* out /* 4096 */);
/* Same for one WIB2 superchunk and register selector: out[(b*12 + f)*16 + lane] (wib2/tpg/FrameExpand.hpp:193-209). */
void oracle_wib2_expand(const uint8_t* superchunk, int sel, uint32_t adc_offset, uint16_t* out /* 1536 */);
/* firwin_int restated (src/wib2/tpg/DesignFIR.cpp:20-68). */
int oracle_firwin_int(int n, double cutoff, int multiplier, int16_t* taps);
void oracle_get_state(const oracle_link* lk, swtpg_channel_state* out /* n_channels */);

#ifdef __cplusplus
}
#endif
#endif
