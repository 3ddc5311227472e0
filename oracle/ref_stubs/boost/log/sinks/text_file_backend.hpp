// oracle test infrastructure: stand-in for <boost/log/sinks/text_file_backend.hpp>; see gsdr_thirdparty_stub.hpp
#pragma once
#include "../../../gsdr_thirdparty_stub.hpp"
