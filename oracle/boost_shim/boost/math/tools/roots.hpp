// TEST INFRASTRUCTURE (oracle/_ref build only). Bracketed Halley iteration standing in for
// boost::math::tools::halley_iterate (cpp/src/rochelobe.cpp:51-62,88-96). Only the companion-star
// geometry uses it; no hot-path output depends on it, so it is not a bit-level restatement.
#pragma once
#include <cmath>
#include <tuple>
namespace boost { namespace math { namespace tools {
template <class F, class T> T halley_iterate(F f, T guess, T min, T max, int digits) {
    const T tol = std::ldexp(T(1), 1 - digits);
    T x = guess;
    for (int it = 0; it < 200; ++it) {
        auto [f0, f1, f2] = f(x);
        if (f0 == 0) return x;
        T denom = 2 * f1 * f1 - f0 * f2;
        T dx;
        if (denom != 0 && std::isfinite(denom)) {
            dx = 2 * f0 * f1 / denom;
        } else if (f1 != 0) {
            dx = f0 / f1;
        } else {
            dx = (x - min > max - x) ? (x - min) / 2 : -(max - x) / 2;
        }
        T xn = x - dx;
        if (xn <= min) { xn = (x + min) / 2; }
        if (xn >= max) { xn = (x + max) / 2; }
        const T step = std::fabs(xn - x);
        x = xn;
        if (step <= std::fabs(x) * tol) break;
    }
    return x;
}
}}}  // namespace boost::math::tools
