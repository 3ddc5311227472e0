// oracle test infrastructure: stand-in for <boost/math/special_functions/round.hpp>; see gsdr_thirdparty_stub.hpp
#pragma once
#include "../../../gsdr_thirdparty_stub.hpp"
