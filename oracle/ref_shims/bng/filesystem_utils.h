/* placeholder for libbng's filesystem_utils.h (absent sibling repo).  TEST INFRASTRUCTURE, own file. */
#pragma once
#include <string>
namespace FSUtils { static inline void make_dir_for_file_w_multiple_attempts(const std::string&) {} }
