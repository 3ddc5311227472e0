// oracle test infrastructure: stand-in for <boost/algorithm/string.hpp>; see gsdr_thirdparty_stub.hpp
#pragma once
#include "../../gsdr_thirdparty_stub.hpp"
