// stand-in for <boost/lambda/bind.hpp>: three-argument bind with placeholders and nested binds, which is the
// form the reference's Xform.hpp uses to compose coordinate transforms
#pragma once
#include <utility>
namespace boost { namespace lambda {
template <int N> struct placeholder {};
static const placeholder<1> _1 = placeholder<1>();
static const placeholder<2> _2 = placeholder<2>();
static const placeholder<3> _3 = placeholder<3>();
template <class F, class A1, class A2, class A3> struct bound3;
namespace detail {
template <class X, class Y, class Z> inline X pick(placeholder<1>, const X &x, const Y &, const Z &) { return x; }
template <class X, class Y, class Z> inline Y pick(placeholder<2>, const X &, const Y &y, const Z &) { return y; }
template <class X, class Y, class Z> inline Z pick(placeholder<3>, const X &, const Y &, const Z &z) { return z; }
template <class F, class A1, class A2, class A3, class X, class Y, class Z>
inline auto pick(const bound3<F, A1, A2, A3> &b, const X &x, const Y &y, const Z &z) -> decltype(b(x, y, z)) { return b(x, y, z); }
}
template <class F, class A1, class A2, class A3>
struct bound3 {
  F f; A1 a1; A2 a2; A3 a3;
  template <class X, class Y, class Z>
  auto operator()(const X &x, const Y &y, const Z &z) const
      -> decltype(f(detail::pick(a1, x, y, z), detail::pick(a2, x, y, z), detail::pick(a3, x, y, z))) {
    return f(detail::pick(a1, x, y, z), detail::pick(a2, x, y, z), detail::pick(a3, x, y, z));
  }
};
template <class F, class A1, class A2, class A3>
inline bound3<F, A1, A2, A3> bind(F f, A1 a1, A2 a2, A3 a3) { return bound3<F, A1, A2, A3>{f, a1, a2, a3}; }
// plain functions (possibly a template-id naming an overload set): resolved against the pointer type
template <class R, class B1, class B2, class B3, class A1, class A2, class A3>
inline bound3<R (*)(B1, B2, B3), A1, A2, A3> bind(R (*f)(B1, B2, B3), A1 a1, A2 a2, A3 a3) {
  return bound3<R (*)(B1, B2, B3), A1, A2, A3>{f, a1, a2, a3};
}
} }  // namespace boost::lambda
