// oracle test infrastructure: stand-in for <boost/log/sources/severity_logger.hpp>; see gsdr_thirdparty_stub.hpp
#pragma once
#include "../../../gsdr_thirdparty_stub.hpp"
