// oracle test infrastructure: stand-in for <uhd/exception.hpp>; see gsdr_thirdparty_stub.hpp
#pragma once
#include "../gsdr_thirdparty_stub.hpp"
