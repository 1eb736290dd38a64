// front.hpp -- internal declarations shared by the host-side sources.
#pragma once
#include <string>

namespace csolve_front {
void set_last_error(const std::string &s);
const char *last_error();
}  // namespace csolve_front
