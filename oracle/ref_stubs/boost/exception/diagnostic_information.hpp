// oracle test infrastructure: stand-in for <boost/exception/diagnostic_information.hpp>; see gsdr_thirdparty_stub.hpp
#pragma once
#include "../../gsdr_thirdparty_stub.hpp"
