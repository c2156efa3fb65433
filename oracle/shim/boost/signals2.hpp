// Minimal stand-in for the subset of boost::signals2 / boost::bind the reference uses
// (signal<void()>, connect, operator(), bind(&T::f, this)). TEST INFRASTRUCTURE ONLY:
// lets the unmodified reference sources under /root/reference compile in an image without Boost.
#pragma once
#include <functional>
#include <vector>
namespace boost {
template <class F, class T> auto bind(F f, T* obj) { return [f, obj]() { (obj->*f)(); }; }
namespace signals2 {
template <class Sig> class signal;
template <> class signal<void()> {
  std::vector<std::function<void()>> slots_;
public:
  template <class F> void connect(F f) { slots_.emplace_back(f); }
  void operator()() { for (auto& s : slots_) s(); }
};
}  // namespace signals2
}  // namespace boost
