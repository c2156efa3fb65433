// Stand-in for boost::iequals / to_lower / to_lower_copy (test infrastructure only).
#pragma once
#include <algorithm>
#include <cctype>
#include <string>
namespace boost {
namespace algorithm {
inline void to_lower(std::string& s) {
  std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return std::tolower(c); });
}
inline std::string to_lower_copy(std::string s) { to_lower(s); return s; }
}  // namespace algorithm
using algorithm::to_lower;
using algorithm::to_lower_copy;
inline bool iequals(const std::string& a, const std::string& b) {
  return algorithm::to_lower_copy(a) == algorithm::to_lower_copy(b);
}
}  // namespace boost
