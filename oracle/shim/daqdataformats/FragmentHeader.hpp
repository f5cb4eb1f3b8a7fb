// Oracle build shim (test infrastructure): FragmentType constants named at
// include/fdreadoutlibs/DUNEWIBEthTypeAdapter.hpp:92 and DUNEWIBSuperChunkTypeAdapter.hpp. Values are not used
// on the hot path.
#pragma once
#include <cstdint>
namespace dunedaq {
namespace daqdataformats {
enum class FragmentType : uint32_t { kUnknown = 0, kProtoWIB = 1, kWIB = 2, kDAPHNE = 3, kTDE_AMC = 4, kWIBEth = 12 };
} // namespace daqdataformats
} // namespace dunedaq
