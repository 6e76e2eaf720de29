// See star.hpp.  Operation order follows the reference so that centres, normals and areas agree to the last bits
// (they feed cosines and a shadow test); the radius solve is a bracketed Halley iteration run to convergence.
#include "star.hpp"

#include <array>
#include <cmath>
#include <tuple>

#include "../../include/freddi_b200.h"

namespace fb200 {
namespace {

// Vec3 with its cached length (cpp/src/geometry.cpp:6-112): the length is computed at construction and only
// rescaled by operator*=; UnitVec3 pins it to exactly 1.
struct V3 {
    double x, y, z, len;
};
inline V3 mk(double x, double y, double z) { return {x, y, z, std::sqrt(x * x + y * y + z * z)}; }
inline V3 add(const V3& a, const V3& b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 sub(const V3& a, const V3& b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 scale(V3 a, double f) { a.x *= f; a.y *= f; a.z *= f; a.len *= f; return a; }
inline V3 div(const V3& a, double f) { return scale(a, 1.0 / f); }       // Vec3::operator/ = *this * (1.0 / factor)
inline V3 cross(const V3& a, const V3& b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline V3 unit(const V3& a) { V3 u = div(a, a.len); u.len = 1.0; return u; }  // UnitVec3(const Vec3&)

typedef std::array<V3, 3> Tri;

std::vector<Tri> unit_sphere_triangles(int lod) {  // cpp/src/geometry.cpp:230-242
    const double X = .525731112119133606, Z = .850650808352039932;
    const double v[12][3] = {{-X, 0.0, Z}, {X, 0.0, Z}, {-X, 0.0, -Z}, {X, 0.0, -Z}, {0.0, Z, X}, {0.0, Z, -X},
                             {0.0, -Z, X}, {0.0, -Z, -X}, {Z, X, 0.0}, {-Z, X, 0.0}, {Z, -X, 0.0}, {-Z, -X, 0.0}};
    const int idx[20][3] = {{0, 1, 4}, {0, 4, 9}, {9, 4, 5}, {4, 8, 5}, {4, 1, 8}, {8, 1, 10}, {8, 10, 3}, {5, 8, 3}, {5, 3, 2}, {2, 3, 7},
                            {7, 3, 10}, {7, 10, 6}, {7, 6, 11}, {11, 6, 0}, {0, 6, 1}, {6, 10, 1}, {9, 11, 0}, {9, 2, 11}, {9, 5, 2}, {7, 11, 2}};
    std::vector<Tri> tris;
    for (const auto& t : idx) tris.push_back({mk(v[t[0]][0], v[t[0]][1], v[t[0]][2]), mk(v[t[1]][0], v[t[1]][1], v[t[1]][2]), mk(v[t[2]][0], v[t[2]][1], v[t[2]][2])});
    for (int i = lod; i != 0; --i) {
        std::vector<Tri> tmp;
        tmp.reserve(tris.size() * 4);
        for (const Tri& T : tris) {  // Triangle::divide + projectedOntoUnitSphere, geometry.cpp:213-227
            const V3 v01 = scale(add(T[0], T[1]), 0.5), v02 = scale(add(T[0], T[2]), 0.5), v12 = scale(add(T[1], T[2]), 0.5);
            const Tri parts[4] = {{T[0], v01, v02}, {T[1], v12, v01}, {T[2], v02, v12}, {v01, v12, v02}};
            for (const Tri& p : parts) tmp.push_back({unit(p[0]), unit(p[1]), unit(p[2])});
        }
        tris.swap(tmp);
    }
    return tris;
}

inline double p2(double x) { return x * x; }
inline double p3(double x) { return x * x * x; }
inline double p4(double x) { const double q = x * x; return q * q; }
inline double p5(double x) { const double q = x * x; return x * q * q; }
inline double p7(double x) { const double q = x * x * x; return x * q * q; }

struct Roche {  // RochePotential, cpp/src/rochelobe.cpp:19-48
    double q;   // mass_ratio
    double omega(double r, double lambda, double nu) const {
        return q / r + (1.0 / std::sqrt(1.0 + p2(r) - 2.0 * r * lambda) - r * lambda) + 0.5 * (1.0 + q) * p2(r) * (1.0 - p2(nu));
    }
    double d1(double r, double lambda, double nu) const {
        return -q / p2(r) - ((r - lambda) * std::pow(1.0 + p2(r) - 2.0 * r * lambda, -1.5) + lambda) + (1.0 + q) * r * (1.0 - p2(nu));
    }
    double d2(double r, double lambda, double nu) const {
        const double rcm2 = 1.0 / std::sqrt(1.0 + p2(r) - 2.0 * r * lambda);
        return 2.0 * q / p3(r) - p3(rcm2) + 3.0 * p2(r - lambda) * p5(rcm2) + (1.0 + q) * (1.0 - p2(nu));
    }
    double d3(double r, double lambda) const {
        const double rcm2 = 1.0 / std::sqrt(1.0 + p2(r) - 2.0 * r * lambda);
        return -6.0 * q / p4(r) + 9.0 * (r - lambda) * p5(rcm2) - 15.0 * p3(r - lambda) * p7(rcm2);
    }
};

// Bracketed Halley iteration to 26 binary digits of step size (boost::math::tools::halley_iterate as called in
// cpp/src/rochelobe.cpp:51-62,88-96).  The iteration converges cubically: once a step is below 2^-25 |x| the root is
// exact to rounding, so the result does not depend on the details of the safeguard.
template <class F>
double halley(F f, double guess, double lo, double hi) {
    const double tol = std::ldexp(1.0, 1 - 26);
    double x = guess;
    for (int it = 0; it < 200; ++it) {
        double f0, f1, f2;
        std::tie(f0, f1, f2) = f(x);
        if (f0 == 0) return x;
        const double denom = 2 * f1 * f1 - f0 * f2;
        double dx;
        if (denom != 0 && std::isfinite(denom)) dx = 2 * f0 * f1 / denom;
        else if (f1 != 0) dx = f0 / f1;
        else dx = (x - lo > hi - x) ? (x - lo) / 2 : -(hi - x) / 2;
        double xn = x - dx;
        if (xn <= lo) xn = (x + lo) / 2;
        if (xn >= hi) xn = (x + hi) / 2;
        const double step = std::fabs(xn - x);
        x = xn;
        if (step <= std::fabs(x) * tol) break;
    }
    return x;
}

double find_r(const Roche& P, double omega_value, double lambda, double nu, double guess, double lo, double hi) {
    return halley([&](double x) { return std::make_tuple(P.omega(x, lambda, nu) - omega_value, P.d1(x, lambda, nu), P.d2(x, lambda, nu)); },
                  guess, lo, hi);
}

}  // namespace

std::vector<double> StarGeometry::packed() const {
    std::vector<double> p;
    p.reserve(size_t(7) * n_tri);
    for (const std::vector<double>* v : {&cx, &cy, &cz, &nx, &ny, &nz, &area}) p.insert(p.end(), v->begin(), v->end());
    return p;
}

int build_star_geometry(double semiaxis, double mass_ratio, double fill, int lod, StarGeometry& out, std::string& err) {
    if (lod < 0 || lod > 6) {
        err = "starlod must be in [0, 6] (20 * 4^starlod triangles)";
        return FB200_ERR_UNSUPPORTED;
    }
    if (!(mass_ratio > 0.) || !(fill > 0.) || !(semiaxis > 0.)) {
        err = "the companion star needs positive Mopt, Mx, period and rochelobefill";
        return FB200_ERR_INVALID_ARGUMENT;
    }
    const Roche P{mass_ratio};
    // CriticalRocheLobe, cpp/src/rochelobe.cpp:65-104
    double lagrangian_point;
    {
        const double q = mass_ratio <= 1.0 ? mass_ratio : 1.0 / mass_ratio;
        double x0 = std::cbrt(q / 3.0) + (0.5 - std::cbrt(1.0 / 3.0)) * q;
        double x1 = 0.5 * x0;
        double x2 = std::min(0.5, 2 * x0);
        if (mass_ratio > 1.0) {
            x0 = 1.0 - x0;
            const double tmp = 1.0 - x1;
            x1 = 1.0 - x2;
            x2 = tmp;
        }
        lagrangian_point = halley([&](double x) { return std::make_tuple(P.d1(x, 1.0, 0.0), P.d2(x, 1.0, 0.0), P.d3(x, 1.0)); }, x0, x1, x2);
    }
    const double omega_crit = P.omega(lagrangian_point, 1.0, 0.0);
    double volume_radius;
    {   // BinaryFunctions::rocheLobeVolumeRadiusSemiaxis (Eggleton 1983), cpp/src/orbit.cpp:19-23
        const double q = std::cbrt(mass_ratio);
        volume_radius = 0.49 * p2(q) / (0.6 * p2(q) + std::log(1. + q));
    }
    const double crit_polar = find_r(P, omega_crit, 0.0, 1.0, volume_radius, 0.5 * volume_radius, lagrangian_point);
    // DimensionlessRocheLobe, cpp/src/rochelobe.cpp:107-126
    const double polar_radius = fill * crit_polar;
    const double omega = P.omega(polar_radius, 0.0, 1.0);
    auto lobe_r = [&](const V3& v) {
        const double lambda = -v.x / v.len, nu = v.z / v.len;
        const double r0 = polar_radius;
        return semiaxis * find_r(P, omega, lambda, nu, r0, 0.5 * r0, std::min(2.0 * r0, lagrangian_point));
    };
    const std::vector<Tri> sphere = unit_sphere_triangles(lod);
    const int n = int(sphere.size());
    out = StarGeometry();
    out.n_tri = n;
    for (std::vector<double>* v : {&out.cx, &out.cy, &out.cz, &out.nx, &out.ny, &out.nz, &out.area}) v->resize(n);
    out.vertices.resize(size_t(9) * n);
    for (int t = 0; t < n; ++t) {  // Star::initializeRocheTriangles (star.cpp:29-40) + Triangle ctor (geometry.cpp:129-136)
        Tri T = sphere[t];
        for (V3& v : T) v = scale(v, lobe_r(v));
        const V3 c = cross(sub(T[1], T[0]), sub(T[2], T[1]));
        const V3 centre = div(add(add(T[0], T[1]), T[2]), 3.0);
        const V3 normal = unit(c);
        out.area[t] = 0.5 * c.len;
        out.cx[t] = centre.x; out.cy[t] = centre.y; out.cz[t] = centre.z;
        out.nx[t] = normal.x; out.ny[t] = normal.y; out.nz[t] = normal.z;
        for (int k = 0; k < 3; ++k) { out.vertices[9 * t + 3 * k] = T[k].x; out.vertices[9 * t + 3 * k + 1] = T[k].y; out.vertices[9 * t + 3 * k + 2] = T[k].z; }
    }
    if (!std::isfinite(out.area[0])) {
        err = "the Roche-lobe surface of the companion star could not be constructed";
        return FB200_ERR_RUNTIME;
    }
    return FB200_OK;
}

}  // namespace fb200
