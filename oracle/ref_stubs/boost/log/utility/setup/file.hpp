// oracle test infrastructure: stand-in for <boost/log/utility/setup/file.hpp>; see gsdr_thirdparty_stub.hpp
#pragma once
#include "../../../../gsdr_thirdparty_stub.hpp"
