// oracle test infrastructure: stand-in for <boost/lockfree/spsc_queue.hpp>; see gsdr_thirdparty_stub.hpp
#pragma once
#include "../../gsdr_thirdparty_stub.hpp"
