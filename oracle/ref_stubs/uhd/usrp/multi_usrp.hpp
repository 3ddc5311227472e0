// oracle test infrastructure: stand-in for <uhd/usrp/multi_usrp.hpp>; see gsdr_thirdparty_stub.hpp
#pragma once
#include "../../gsdr_thirdparty_stub.hpp"
