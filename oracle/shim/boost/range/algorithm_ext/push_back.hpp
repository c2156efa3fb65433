// Stand-in for boost::push_back(container, range) (test infrastructure only).
#pragma once
namespace boost { template <class C, class R> void push_back(C& c, const R& r) { for (auto x : r) c.push_back(x); } }
