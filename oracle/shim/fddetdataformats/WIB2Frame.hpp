// Oracle build shim (test infrastructure): restated layout of fddetdataformats::WIB2Frame (dependency NOT under
// /root/reference). What the reference pins: 12 * sizeof == 5664, i.e. sizeof == 472
// (include/fdreadoutlibs/DUNEWIBSuperChunkTypeAdapter.hpp:100); adc_words is uint32_t[112] holding 256 x 14-bit
// channels, channel c at bits [14c, 14c+14) (include/fdreadoutlibs/wib2/tpg/FrameExpand.hpp:193-209,
// test/apps/wib2_test_bench.cxx:233-254); header.timestamp_1/_2 are the low/high 32 bits of the timestamp and
// header.{crate,slot,link,detector_id} exist (DUNEWIBSuperChunkTypeAdapter.hpp:40-66,
// src/wib2/WIB2FrameProcessor.cpp:362). The header LENGTH (20 B here) is restated from memory and unpinned, so the
// B200 unpacker takes the ADC byte offset as a parameter (include/swtpg.h: swtpg_config.wib2_adc_offset).
#pragma once
#include <cstdint>
namespace dunedaq {
namespace fddetdataformats {
class WIB2Frame
{
public:
  typedef uint32_t word_t;
  static constexpr int s_bits_per_adc = 14;
  static constexpr int s_bits_per_word = 8 * sizeof(word_t);
  static constexpr int s_num_ch_per_frame = 256;
  static constexpr int s_num_adc_words = s_num_ch_per_frame * s_bits_per_adc / s_bits_per_word; // 112
  struct Header
  {
    word_t version : 6, detector_id : 6, crate : 10, slot : 4, link : 6;
    word_t timestamp_1;
    word_t timestamp_2;
    word_t colddata_timestamp_id : 1, femb_valid : 2, link_mask : 8, lock_output_status : 1, reserved : 20;
    word_t femb_pulser_frame_bits : 8, femb_sync_flags : 8, colddata_timestamp_0 : 15, reserved_2 : 1;
  };
  struct Trailer
  {
    word_t flex_bits : 16, ws : 1, psr_cal : 4, ready : 1, context_code : 8, reserved : 2;
  };
  Header header;
  word_t adc_words[s_num_adc_words];
  Trailer trailer;

  uint16_t get_adc(int i) const
  {
    const int bit = s_bits_per_adc * i;
    const int w = bit / s_bits_per_word, off = bit % s_bits_per_word;
    uint64_t v = adc_words[w] >> off;
    if (off + s_bits_per_adc > s_bits_per_word)
      v |= uint64_t(adc_words[w + 1]) << (s_bits_per_word - off);
    return static_cast<uint16_t>(v & 0x3fffu);
  }
  void set_adc(int i, uint16_t val)
  {
    const int bit = s_bits_per_adc * i;
    const int w = bit / s_bits_per_word, off = bit % s_bits_per_word;
    const uint32_t v = val & 0x3fffu;
    adc_words[w] = (adc_words[w] & ~(uint32_t(0x3fff) << off)) | (v << off);
    if (off + s_bits_per_adc > s_bits_per_word) {
      const int done = s_bits_per_word - off;
      const uint32_t mask = (uint32_t(1) << (s_bits_per_adc - done)) - 1;
      adc_words[w + 1] = (adc_words[w + 1] & ~mask) | (v >> done);
    }
  }
  uint64_t get_timestamp() const { return uint64_t(header.timestamp_1) | (uint64_t(header.timestamp_2) << 32); }
  void set_timestamp(uint64_t ts)
  {
    header.timestamp_1 = static_cast<word_t>(ts);
    header.timestamp_2 = static_cast<word_t>(ts >> 32);
  }
};
static_assert(sizeof(WIB2Frame) == 472, "WIB2Frame shim must be 472 bytes");
} // namespace fddetdataformats
} // namespace dunedaq
