// oracle test infrastructure: stand-in for <uhd/utils/static.hpp>; see gsdr_thirdparty_stub.hpp
#pragma once
#include "../../gsdr_thirdparty_stub.hpp"
