// common.hpp — error reporting shared by the host translation units.
#pragma once
#include <string>
#include "../../include/eirgrid_b200.h"

// stores the message for eg_last_error() and returns `code`
int eg_fail(int code, const std::string& message);
