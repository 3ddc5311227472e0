// oracle test infrastructure: stand-in for <boost/log/utility/setup/common_attributes.hpp>; see gsdr_thirdparty_stub.hpp
#pragma once
#include "../../../../gsdr_thirdparty_stub.hpp"
