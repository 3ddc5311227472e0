// oracle test infrastructure: stand-in for <boost/date_time/posix_time/posix_time.hpp>; see gsdr_thirdparty_stub.hpp
#pragma once
#include "../../../gsdr_thirdparty_stub.hpp"
