// Oracle build shim (test infrastructure): the one enum of daqdataformats::SourceID the type adapters name
// (include/fdreadoutlibs/DUNEWIBEthTypeAdapter.hpp:91). daqdataformats is not under /root/reference.
#pragma once
#include <cstdint>
namespace dunedaq {
namespace daqdataformats {
using timestamp_t = uint64_t;
struct SourceID
{
  enum class Subsystem : uint16_t { kUnknown = 0, kDetectorReadout = 1, kHwSignalsInterface = 2, kTrigger = 3, kTRBuilder = 4 };
  Subsystem subsystem{ Subsystem::kUnknown };
  uint32_t id{ 0 };
};
} // namespace daqdataformats
} // namespace dunedaq
