// Oracle build shim (test infrastructure). The reference includes <boost/thread/future.hpp> at
// include/fdreadoutlibs/wibeth/tpg/ProcessingInfo.hpp:14-15 but uses nothing from it. Intentionally empty.
#pragma once
