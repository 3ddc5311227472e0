// oracle test infrastructure: stand-in for <boost/asio/use_future.hpp>; see gsdr_thirdparty_stub.hpp
#pragma once
#include "../../gsdr_thirdparty_stub.hpp"
