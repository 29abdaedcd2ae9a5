// shadows the Parthenon header of the same name for the oracle/_ref build (see apk_ref_stub.hpp)
#include "apk_ref_stub.hpp"
