// Oracle build shim (test infrastructure): restated layout of fddetdataformats::WIBEthFrame, a dependency that is
// NOT under /root/reference (version unpinned, CMakeLists.txt:17). What the reference pins about it:
//   * sizeof == 7200 (include/fdreadoutlibs/DUNEWIBEthTypeAdapter.hpp:98)
//   * adc_words is word_t[64][14] with word_t = uint64_t, one row per time sample
//     (include/fdreadoutlibs/wibeth/tpg/FrameExpand.hpp:199-205,215)
//   * channel c of a row sits at bits [14c, 14c+14) (unittest/WIBEthFrameExpansion_test.cxx:105-150 via set_adc)
//   * the timestamp is the 2nd 64-bit word of the frame (docs/README.md:81)
// The bit positions of det_id/crate_id/slot_id/stream_id/seq_id inside word 0 are restated from memory of
// fddetdataformats and are NOT pinned by anything in the reference ("parity unpinned" for those fields only).
#pragma once
#include <cstdint>
#include <cstring>
namespace dunedaq {
namespace fddetdataformats {
struct DAQEthHeader
{
  uint64_t version : 6, det_id : 6, crate_id : 10, slot_id : 4, stream_id : 8, reserved : 6, seq_id : 12, block_length : 12;
  uint64_t timestamp;
};
struct WIBEthHeader
{
  uint64_t word0;
  uint64_t word1;
};
class WIBEthFrame
{
public:
  typedef uint64_t word_t;
  static constexpr int s_bits_per_adc = 14;
  static constexpr int s_bits_per_word = 8 * sizeof(word_t);
  static constexpr int s_time_samples_per_frame = 64;
  static constexpr int s_channels_per_half_femb = 64;
  static constexpr int s_num_channels = 64;
  static constexpr int s_num_adc_words_per_ts = s_num_channels * s_bits_per_adc / s_bits_per_word; // 14

  DAQEthHeader daq_header;
  WIBEthHeader header;
  word_t adc_words[s_time_samples_per_frame][s_num_adc_words_per_ts];

  uint16_t get_adc(int i, int sample = 0) const
  {
    const int bit = s_bits_per_adc * i;
    const int w = bit / s_bits_per_word, off = bit % s_bits_per_word;
    uint64_t v = adc_words[sample][w] >> off;
    if (off + s_bits_per_adc > s_bits_per_word)
      v |= adc_words[sample][w + 1] << (s_bits_per_word - off);
    return static_cast<uint16_t>(v & 0x3fffu);
  }
  void set_adc(int i, int sample, uint16_t val)
  {
    const int bit = s_bits_per_adc * i;
    const int w = bit / s_bits_per_word, off = bit % s_bits_per_word;
    const uint64_t v = val & 0x3fffu;
    adc_words[sample][w] = (adc_words[sample][w] & ~(uint64_t(0x3fff) << off)) | (v << off);
    if (off + s_bits_per_adc > s_bits_per_word) {
      const int done = s_bits_per_word - off;
      const uint64_t mask = (uint64_t(1) << (s_bits_per_adc - done)) - 1;
      adc_words[sample][w + 1] = (adc_words[sample][w + 1] & ~mask) | (v >> done);
    }
  }
  uint64_t get_timestamp() const { return daq_header.timestamp; }
  void set_timestamp(uint64_t ts) { daq_header.timestamp = ts; }
};
static_assert(sizeof(WIBEthFrame) == 7200, "WIBEthFrame shim must be 7200 bytes");
} // namespace fddetdataformats
} // namespace dunedaq
