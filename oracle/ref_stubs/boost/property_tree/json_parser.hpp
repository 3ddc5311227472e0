// oracle test infrastructure: stand-in for <boost/property_tree/json_parser.hpp>; see gsdr_thirdparty_stub.hpp
#pragma once
#include "../../gsdr_thirdparty_stub.hpp"
