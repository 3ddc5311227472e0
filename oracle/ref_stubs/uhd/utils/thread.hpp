// oracle test infrastructure: stand-in for <uhd/utils/thread.hpp>; see gsdr_thirdparty_stub.hpp
#pragma once
#include "../../gsdr_thirdparty_stub.hpp"
