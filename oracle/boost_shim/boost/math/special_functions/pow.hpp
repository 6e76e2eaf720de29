// TEST INFRASTRUCTURE (oracle/_ref build only). Restatement of boost::math::pow<N>(x): compile-time
// exponentiation by squaring with Boost's multiplication order (even N: p=pow<N/2>, p*p; odd N: x*p*p;
// N<0: 1/pow<-N>; integral arguments promote to double).  Parity with real Boost is NOT verifiable
// offline ("parity unpinned" at the last-bit level, see oracle/README.md).
#pragma once
#include <type_traits>
namespace boost { namespace math {
namespace detail_shim {
template <int N, int M = N % 2> struct positive_power {
    template <class T> static constexpr T result(T base) {
        T power = positive_power<N / 2>::result(base);
        return power * power;
    }
};
template <int N> struct positive_power<N, 1> {
    template <class T> static constexpr T result(T base) {
        T power = positive_power<N / 2>::result(base);
        return base * power * power;
    }
};
template <> struct positive_power<1, 1> {
    template <class T> static constexpr T result(T base) { return base; }
};
template <class T> using promote_t = std::conditional_t<std::is_integral<T>::value, double, T>;
}  // namespace detail_shim
template <int N, class T> constexpr detail_shim::promote_t<T> pow(T x) {
    using R = detail_shim::promote_t<T>;
    if constexpr (N == 0) {
        return R(1);
    } else if constexpr (N > 0) {
        return detail_shim::positive_power<N>::result(static_cast<R>(x));
    } else {
        return R(1) / detail_shim::positive_power<-N>::result(static_cast<R>(x));
    }
}
}}  // namespace boost::math
