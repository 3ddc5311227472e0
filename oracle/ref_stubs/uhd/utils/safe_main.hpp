// oracle test infrastructure: stand-in for <uhd/utils/safe_main.hpp>; see gsdr_thirdparty_stub.hpp
#pragma once
#include "../../gsdr_thirdparty_stub.hpp"
