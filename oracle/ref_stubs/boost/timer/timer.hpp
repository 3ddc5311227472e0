// oracle test infrastructure: stand-in for <boost/timer/timer.hpp>; see gsdr_thirdparty_stub.hpp
#pragma once
#include "../../gsdr_thirdparty_stub.hpp"
