#include "../../gsdr_more_stubs.hpp"
