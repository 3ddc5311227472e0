// oracle test infrastructure: stand-in for <boost/log/core.hpp>; see gsdr_thirdparty_stub.hpp
#pragma once
#include "../../gsdr_thirdparty_stub.hpp"
