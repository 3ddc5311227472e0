// oracle test infrastructure: stand-in for <boost/atomic.hpp>; see gsdr_thirdparty_stub.hpp
#pragma once
#include "../gsdr_thirdparty_stub.hpp"
