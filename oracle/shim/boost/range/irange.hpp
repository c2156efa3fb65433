// Stand-in for boost::irange(int,int) (test infrastructure only).
#pragma once
#include <vector>
namespace boost {
inline std::vector<int> irange(int a, int b) { std::vector<int> v; for (int i = a; i < b; ++i) v.push_back(i); return v; }
}
