// Oracle build shim (test infrastructure). Shadows the reference header of the same name, which
// include/fdreadoutlibs/wib2/tpg/FrameExpand.hpp:13 pulls in but does not use; the real one needs
// fddetdataformats/WIBFrame.hpp (ProtoWIB), which is outside the hot path. Intentionally empty.
#pragma once
