// oracle test infrastructure: stand-in for <boost/format.hpp>; see gsdr_thirdparty_stub.hpp
#pragma once
#include "../gsdr_thirdparty_stub.hpp"
