// Host-side argument resolution (see resolve.hpp).  Every function names the reference code it restates.
#include "resolve.hpp"

#include <cmath>
#include <cstring>
#include <functional>
#include <limits>
#include <optional>

#include "fb200_math.h"

namespace fb200 {
namespace {

struct Fail {
    int code;
    std::string msg;
};
[[noreturn]] void fail(int code, const std::string& msg) { throw Fail{code, msg}; }

std::optional<double> opt(double v) { return std::isnan(v) ? std::optional<double>() : std::optional<double>(v); }

// ---------------------------------------------------------------------------------------------
// Cash-Karp RK5 with a constant step, the way boost::numeric::odeint::integrate_adaptive drives a
// plain (uncontrolled) runge_kutta_cash_karp54 stepper: the reference calls it like that for the
// initial-condition normalisation integrals (cpp/src/arguments.cpp:71-82,109-120,141-152,167-178,
// 215-227,322-335).  The integrand does not depend on the state, the state is a scalar; stage sums
// are `1*x + (a*dt)*k...` left to right as odeint's generic Runge-Kutta algorithm forms them.
// ---------------------------------------------------------------------------------------------
double ck5_const_step(const std::function<double(double)>& f, double t0, double t1, double dt) {
    static const double c[6] = {0., 1. / 5., 3. / 10., 3. / 5., 1., 7. / 8.};
    static const double b[6] = {37. / 378., 0., 250. / 621., 125. / 594., 0., 512. / 1771.};
    const double eps = std::numeric_limits<double>::epsilon();
    double x = 0.;
    auto do_step = [&](double t, double step) {
        double k[6];
        for (int s = 0; s < 6; ++s) k[s] = f(t + c[s] * step);  // k[0] at t (c[0] = 0)
        x = 1 * x + (b[0] * step) * k[0] + (b[1] * step) * k[1] + (b[2] * step) * k[2] + (b[3] * step) * k[3] +
            (b[4] * step) * k[4] + (b[5] * step) * k[5];
    };
    double time = t0;
    int step = 0;
    while ((time + dt) - t1 <= eps) {  // less_eq_with_sign(time + dt, t1, dt), dt > 0
        do_step(time, dt);
        ++step;
        time = t0 + static_cast<double>(step) * dt;
    }
    const double end = t0 + dt * static_cast<size_t>(step);
    if (t1 - end > eps) {  // closing partial step
        do_step(end, t1 - end);
    }
    return x;
}

// cpp/src/orbit.cpp:9-23
double r_kerrISCORg(double kerr) {
    const double Z1 = 1. + std::cbrt((1. - p2(kerr))) * (std::cbrt((1. + kerr)) + std::cbrt((1. - kerr)));
    const double Z2 = std::sqrt(3. * p2(kerr) + p2(Z1));
    return 3. + Z2 - std::sqrt((3. - Z1) * (3. + Z1 + 2. * Z2));
}
double efficiency_of_accretion(double kerr) { return 1. - std::sqrt(1. - 2. / 3. / r_kerrISCORg(kerr)); }
// cpp/include/unit_transformation.hpp:15-17, cpp/include/orbit.hpp:17
double r_kerrISCO(double Mx, double kerr) { return r_kerrISCORg(kerr) * FB_G_GRAV * Mx / (FB_C_LIGHT * FB_C_LIGHT); }
// cpp/src/orbit.cpp:19-23, cpp/include/orbit.hpp:24-32
double rocheLobeVolumeRadiusSemiaxis(double mass_ratio) {
    const double q = std::cbrt(mass_ratio);
    return 0.49 * p2(q) / (0.6 * p2(q) + std::log(1. + q));
}
double semiaxis(double total_mass, double period) {
    return std::cbrt(FB_G_GRAV * total_mass * p2(period) / (4. * p2(M_PI)));
}

// cpp/src/opacity_related.cpp:4-93
struct Oprel {
    double m, n, D, Height_coef, Height_exp_F, Height_exp_R, a0, a1, a2, k, l;
    double f_F(double xi) const { return a0 * xi + a1 * pow(xi, k) + a2 * pow(xi, l); }
};
Oprel make_oprel(int opacity, double Mx, double alpha, double mu) {
    Oprel o{};
    const double GM = FB_G_GRAV * Mx;
    if (opacity == FB200_OPACITY_KRAMERS) {
        o.m = 0.3;
        o.n = 0.8;
        const double varkappa0 = 5e24;
        const double Pi1 = 6.31217, Pi2 = 0.51523, Pi3 = 1.13050, Pi4 = 0.39842;
        const double Pi_Sigma = pow(Pi1, 0.05) * pow(Pi2, 0.1) * pow(Pi3, 0.8) * pow(Pi4, 0.1);
        const double Pi_Height = pow(Pi1, 19. / 40.) * pow(Pi2, -0.05) * pow(Pi3, 0.1) * pow(Pi4, -0.05);
        o.D = 2.41e34 * pow(alpha, 0.8) * (Mx / FB_SOLAR_MASS) * pow(mu / 0.5, -0.75) / Pi_Sigma * pow(varkappa0 / 1e22, 0.1);
        o.Height_exp_F = 3. / 20.;
        o.Height_exp_R = 1. / 8.;
        o.Height_coef = 0.020 * pow(1e17, -o.Height_exp_F) * pow(GM, -o.Height_exp_F / 2.) * pow(1e10, -o.Height_exp_F / 2.) *
                        pow(alpha, -0.1) * pow(Mx / FB_SOLAR_MASS, -3. / 8.) * pow(mu / 0.6, -3. / 8.) * Pi_Height *
                        pow(varkappa0 / 5e24, 0.05);
        o.a0 = 1.430;
        o.a1 = -0.46;
        o.a2 = 0.03;
        o.k = 3.5;
        o.l = 6.0;
    } else if (opacity == FB200_OPACITY_OPAL) {
        o.m = 1. / 3.;
        o.n = 1.;
        const double varkappa0 = 1.5e20;
        const double Pi1 = 5.53450, Pi2 = 0.55220, Pi3 = 1.14712, Pi4 = 0.44777;
        const double Pi_Sigma = pow(Pi1, 1. / 18.) * pow(Pi2, 1. / 9.) * pow(Pi3, 7. / 9.) * pow(Pi4, 1. / 9.);
        const double Pi_Height = pow(Pi1, 17. / 36.) * pow(Pi2, -1. / 18.) * pow(Pi3, 1. / 9.) * pow(Pi4, -1. / 18.);
        o.D = 2.12e37 * pow(alpha, 7. / 9.) * pow(Mx / FB_SOLAR_MASS, 10. / 9.) * pow(mu / 0.5, -13. / 18.) / Pi_Sigma *
              pow(varkappa0 / 1e22, 1. / 9.);
        o.Height_exp_F = 1. / 6.;
        o.Height_exp_R = 1. / 12.;
        o.Height_coef = 0.021 * pow(1e17, -o.Height_exp_F) * pow(GM, -o.Height_exp_F / 2.) * pow(1e10, -o.Height_exp_F / 2.) *
                        pow(alpha, -1. / 9.) * pow(Mx / FB_SOLAR_MASS, -13. / 36.) * pow(mu / 0.6, -13. / 36.) * Pi_Height *
                        pow(varkappa0 / 1.5e20, 1. / 18.);
        o.a0 = 1.3998;
        o.a1 = -0.425;
        o.a2 = 0.0249;
        o.k = 11. / 3.;
        o.l = 19. / 3.;
    } else {
        fail(FB200_ERR_INVALID_ARGUMENT, "Wrong opacity");  // std::invalid_argument(opacity_type)
    }
    return o;
}

// Per-thread one-entry memos of the resolver.  In a parameter sweep most models share the grid (it depends on Mx, rin, rout only) and
// the shape of the initial condition; the normalisation integral (600 integrand evaluations of four pow each), the grid (one pow per
// point) and f_F on the grid (two pow per point) are 90 % of resolving a quasistat model.  A hit returns the very doubles the same code
// computed for the same inputs, so results are bit-identical with and without the memo.
struct QuasistatIntegralMemo {
    double key[9];
    double integral;
    bool valid = false;
};
struct GridMemo {
    double h_in = 0., h_out = 0.;
    size_t Nx = 0;
    int scale = -1;
    std::vector<double> h;
};
struct FfMemo {
    double key[10];  // h.front(), h[1], h[N/2], h.back(), N, a0, a1, a2, k, l (the grids are log or linear: these points fix them)
    std::vector<double> fF;
    bool valid = false;
};

// cpp/src/arguments.cpp:321-335
double quasistat_coeff(double h_in, double h_out, const Oprel& oprel) {
    const double x0 = h_in / (h_out - h_in);
    const double x1 = h_in / h_out;
    static thread_local QuasistatIntegralMemo memo;
    const double key[9] = {h_in, h_out, oprel.m, oprel.n, oprel.a0, oprel.a1, oprel.a2, oprel.k, oprel.l};
    if (!(memo.valid && std::memcmp(memo.key, key, sizeof(key)) == 0)) {
        memo.integral = ck5_const_step(
            [&](double x) {
                return pow(oprel.f_F(x * (1. - x1) + x1) * x / (x * (1. - x1) + x1), 1. - oprel.m) * pow(x + x0, oprel.n);
            },
            0., 1., 0.01);
        std::memcpy(memo.key, key, sizeof(key));
        memo.valid = true;
    }
    const double integral = memo.integral;
    return std::pow(h_out - h_in, oprel.n + 1.) * integral / ((1. - oprel.m) * oprel.D);
}

// SibgatullinSunyaev2000Geometry, cpp/src/ns/ns_arguments.cpp:55-69
double ss2000_radiusNS(double freqx) {
    const double f = freqx / 1000.0;
    const double r_km = 12.44 - 3.061 * f + 0.843 * p2(f) + 0.6 * p3(f) + 1.56 * p4(f);
    return 1e5 * r_km;
}
double ss2000_radiusISCO(double freqx) {
    const double f = freqx / 1000.0;
    const double isco_minus_ns_km = 1.44 - 3.061 * f + 0.843 * p2(f) + 0.6 * p3(f) - 0.22 * p4(f);
    return 1e5 * isco_minus_ns_km + ss2000_radiusNS(freqx);
}

struct InitialF {
    int kind;  // FB200_IC_*
    double F0, Mdisk0, Mdot0, powerorder, gaussmu, gausssigma, h_in_ns;
};

// cpp/src/arguments.cpp:57-269
InitialF normalise_initial_F(const Oprel& oprel, int initialcond, double h_in, double h_out, std::optional<double> F0,
                             std::optional<double> Mdisk0, std::optional<double> Mdot0, std::optional<double> powerorder,
                             std::optional<double> gaussmu, std::optional<double> gausssigma) {
    if (!F0 && !Mdisk0 && !Mdot0) fail(FB200_ERR_RUNTIME, "One of F0, Mdisk0 or Mdot0 must be specified");
    InitialF r{};
    r.kind = initialcond;
    const double x0 = h_in / (h_out - h_in);
    auto power_integral = [&](double a, double b) {
        return ck5_const_step([a, b, x0](double x) { return pow(x, a) * pow(x + x0, b); }, 0., 1., 0.01);
    };
    if (initialcond == FB200_IC_LINEARF) {
        const double integral = power_integral(1. - oprel.m, oprel.n);
        const double coeff = std::pow(h_out - h_in, oprel.n + 1.) * integral / ((1. - oprel.m) * oprel.D);
        if (Mdot0) {
            F0 = *Mdot0 * (h_out - h_in);
            Mdisk0 = std::pow(*F0, 1. - oprel.m) * coeff;
        } else if (Mdisk0) {
            F0 = std::pow(*Mdisk0 / coeff, 1. / (1. - oprel.m));
            Mdot0 = *F0 / (h_out - h_in);
        } else {
            Mdot0 = *F0 / (h_out - h_in);
            Mdisk0 = std::pow(*F0, 1. - oprel.m) * coeff;
        }
        r.kind = FB200_IC_POWERF;  // InitialFLinearF is InitialFPowerF with powerorder 1 (arguments.hpp:126-130)
        r.powerorder = 1.0;
    } else if (initialcond == FB200_IC_POWERF) {
        if (!powerorder) fail(FB200_ERR_RUNTIME, "powerorder must be specified for the case of initialcond=powerF");
        if (Mdot0) fail(FB200_ERR_RUNTIME, "Set Mdisk0 or F0 instead of Mdot0 if initialcond=powerF");
        const double integral = power_integral((1. - oprel.m) * *powerorder, oprel.n);
        const double coeff = std::pow(h_out - h_in, oprel.n + 1.) * integral / ((1. - oprel.m) * oprel.D);
        if (Mdisk0) {
            F0 = std::pow(*Mdisk0 / coeff, 1. / (1. - oprel.m));
        } else {
            Mdisk0 = std::pow(*F0, 1. - oprel.m) * coeff;
        }
        Mdot0 = 0.0;
        r.powerorder = *powerorder;
    } else if (initialcond == FB200_IC_POWERSIGMA) {
        if (!powerorder) fail(FB200_ERR_RUNTIME, "powerorder must be specified for the case of initialcond=powerSigma");
        if (Mdot0) fail(FB200_ERR_RUNTIME, "Set Mdisk0 or F0 instead of Mdot0 if initialcond=powerSigma");
        const double integral = power_integral(*powerorder, 3.);
        const double coeff = p4(h_out - h_in) * integral / ((1. - oprel.m) * oprel.D * std::pow(h_out, 3. - oprel.n));
        if (Mdisk0) {
            F0 = std::pow(*Mdisk0 / coeff, 1. / (1. - oprel.m));
        } else {
            Mdisk0 = std::pow(*F0, 1. - oprel.m) * coeff;
        }
        Mdot0 = 0.0;
        r.powerorder = *powerorder;
    } else if (initialcond == FB200_IC_SINEF) {
        const double a = 1. - oprel.m, b = oprel.n;
        const double integral =
            ck5_const_step([a, b, x0](double x) { return pow(sin(x * M_PI_2), a) * pow(x + x0, b); }, 0., 1., 0.01);
        const double coeff = pow(h_out - h_in, oprel.n + 1.) * integral / ((1. - oprel.m) * oprel.D);
        if (Mdot0) {
            F0 = *Mdot0 * (h_out - h_in) * 2. / M_PI;
            Mdisk0 = std::pow(*F0, 1. - oprel.m) * coeff;
        } else if (Mdisk0) {
            F0 = std::pow(*Mdisk0 / coeff, 1. / (1. - oprel.m));
            Mdot0 = *F0 * M_PI / (2. * (h_out - h_in));
        } else {
            Mdot0 = *F0 * M_PI / (2. * (h_out - h_in));
            Mdisk0 = std::pow(*F0, 1. - oprel.m) * coeff;
        }
    } else if (initialcond == FB200_IC_GAUSSF) {
        if (!gaussmu) fail(FB200_ERR_RUNTIME, "gaussmu must be specified for the case of initialcond=gaussF");
        if (!gausssigma) fail(FB200_ERR_RUNTIME, "gausssigma must be specified for the case of initialcond=gaussF");
        if (*gaussmu <= 0. or *gaussmu > 1.) fail(FB200_ERR_RUNTIME, "gaussmu value should be in [0..1]");
        if (Mdot0)
            fail(FB200_ERR_RUNTIME,
                 "Setting Mdot if initialcond=gaussF produces unstable results and it isn't motivated physically. Set F0 or "
                 "Mdisk0 instead");
        const double a = 1. - oprel.m, b = oprel.n, gm = *gaussmu, gs = *gausssigma;
        const double integral = ck5_const_step(
            [a, b, x0, gm, gs](double x) { return exp(-p2(x - gm) * a / (2. * p2(gs))) * pow(x + x0, b); }, 0., 1., 0.01);
        const double coeff = std::pow(h_out - h_in, oprel.n + 1.) * integral / ((1. - oprel.m) * oprel.D);
        if (Mdisk0) {
            F0 = std::pow(*Mdisk0 / coeff, 1. / (1. - oprel.m));
        } else {
            Mdisk0 = std::pow(*F0, 1. - oprel.m) * coeff;
        }
        Mdot0 = 0.0;
        r.gaussmu = gm;
        r.gausssigma = gs;
    } else if (initialcond == FB200_IC_QUASISTAT) {
        const double coeff = quasistat_coeff(h_in, h_out, oprel);
        if (Mdot0) {
            F0 = *Mdot0 * (h_out - h_in) / h_out * h_in / oprel.f_F(h_in / h_out);
            Mdisk0 = std::pow(*F0, 1. - oprel.m) * coeff;
        } else if (Mdisk0) {
            F0 = std::pow(*Mdisk0 / coeff, 1. / (1. - oprel.m));
            Mdot0 = *F0 * h_out / (h_out - h_in) * oprel.f_F(h_in / h_out) / h_in;
        } else {
            Mdisk0 = std::pow(*F0, 1. - oprel.m) * coeff;
            Mdot0 = *F0 * h_out / (h_out - h_in) * oprel.f_F(h_in / h_out) / h_in;
        }
    } else {
        fail(FB200_ERR_RUNTIME, "Invalid value of initialcond");
    }
    r.F0 = *F0;
    r.Mdisk0 = *Mdisk0;
    r.Mdot0 = *Mdot0;
    return r;
}

// InitialF*::operator() and ::first, cpp/src/arguments.cpp:273-351, cpp/src/ns/ns_arguments.cpp:174-189
void initial_F(const InitialF& ic, const Oprel& oprel, const std::vector<double>& h, std::vector<double>& F, size_t& first) {
    const size_t N = h.size();
    F.assign(N, 0.0);
    first = 0;
    switch (ic.kind) {
        case FB200_IC_POWERF:
            for (size_t i = 0; i < N; ++i) F[i] = ic.F0 * std::pow((h[i] - h.front()) / (h.back() - h.front()), ic.powerorder);
            break;
        case FB200_IC_POWERSIGMA:
            for (size_t i = 0; i < N; ++i) {
                const double Sigma_to_Sigmaout = std::pow((h[i] - h.front()) / (h.back() - h.front()), ic.powerorder);
                F[i] = ic.F0 * std::pow(h[i] / h.back(), (3. - oprel.n) / (1. - oprel.m)) *
                       std::pow(Sigma_to_Sigmaout, 1. / (1. - oprel.m));
            }
            break;
        case FB200_IC_SINEF:
            for (size_t i = 0; i < N; ++i) F[i] = ic.F0 * std::sin((h[i] - h.front()) / (h.back() - h.front()) * M_PI_2);
            break;
        case FB200_IC_QUASISTAT: {
            static thread_local FfMemo memo;
            const double key[10] = {h.front(), h[N > 1 ? 1 : 0], h[N / 2], h.back(), double(N), oprel.a0, oprel.a1, oprel.a2, oprel.k, oprel.l};
            if (!(memo.valid && std::memcmp(memo.key, key, sizeof(key)) == 0)) {
                memo.fF.resize(N);
                for (size_t i = 0; i < N; ++i) memo.fF[i] = oprel.f_F(h[i] / h.back());
                std::memcpy(memo.key, key, sizeof(key));
                memo.valid = true;
            }
            for (size_t i = 0; i < N; ++i) F[i] = ic.F0 * memo.fF[i] * (1. - h.front() / h[i]) / (1. - h.front() / h.back());
            break;
        }
        case FB200_IC_GAUSSF:
            for (size_t i = 0; i < N; ++i) {
                const double xi = (h[i] - h.front()) / (h.back() - h.front());
                F[i] = ic.F0 * std::exp(-p2(xi - ic.gaussmu) / (2. * p2(ic.gausssigma)));
            }
            if ((F[1] == 0.0) || (F[N - 1] == 0.0)) {
                for (size_t i = 0; i < N; ++i) {
                    const double xi = (h[i] - h.front()) / (h.back() - h.front());
                    F[i] += 1e-100 * ic.F0 * xi;
                }
            }
            break;
        case FB200_IC_QUASISTAT_NS: {
            size_t i0 = 0;
            for (;; ++i0) {
                if (i0 >= N) fail(FB200_ERR_LOGIC, "quasistat-ns: Alfven radius is beyond the outer radius");  // h.at(i) -> std::out_of_range
                if (!(h[i0] < ic.h_in_ns)) break;
            }
            first = i0;
            for (size_t i = first; i < N; ++i) {
                const double xi = h[i] / h.back();
                F[i] = ic.F0 * oprel.f_F(xi) * (1. - ic.h_in_ns / h[i]) / (1. - ic.h_in_ns / h.back());
            }
            break;
        }
        default:
            fail(FB200_ERR_RUNTIME, "Invalid value of initialcond");
    }
}

int resolve_impl(const fb200_args_t& a, Resolved& out) {
    fb200_model_dev_t& d = out.dev;
    std::memset(&d, 0, sizeof(d));
    std::memset(&out.info, 0, sizeof(out.info));
    const bool is_ns = a.is_ns != 0;
    d.is_ns = is_ns ? 1 : 0;

    // ---- NeutronStarArguments, cpp/include/ns/ns_arguments.hpp:49-65, cpp/src/ns/ns_arguments.cpp:17-48 ----
    double freqx = 0, Rx = 0, mu_magn = 0;
    if (is_ns) {
        const auto ofreq = opt(a.freqx), oRx = opt(a.Rx);
        if (ofreq) {
            freqx = *ofreq;
        } else {
            if (a.nsprop == FB200_NSPROP_DUMMY) freqx = 0.;
            else if (a.nsprop == FB200_NSPROP_SIBSUN2000) fail(FB200_ERR_RUNTIME, "freqx must be specified for nsprop=sibgatullinsunyaev2000");
            else if (a.nsprop == FB200_NSPROP_NEWT) fail(FB200_ERR_RUNTIME, "freqx must be specified for nsprop=newt");
            else fail(FB200_ERR_INVALID_ARGUMENT, "Wrong nsprop value");
        }
        if (oRx) {
            Rx = *oRx;
        } else {
            if (a.nsprop == FB200_NSPROP_DUMMY) Rx = 1e6;
            else if (a.nsprop == FB200_NSPROP_SIBSUN2000 || a.nsprop == FB200_NSPROP_NEWT) Rx = ss2000_radiusNS(freqx);
            else fail(FB200_ERR_INVALID_ARGUMENT, "Wrong nsprop value");
        }
        mu_magn = 0.5 * a.Bx * p3(Rx);
    }

    // ---- BasicDiskBinaryArguments, cpp/include/arguments.hpp:77-96; NS: cpp/src/ns/ns_arguments.cpp:72-109 ----
    const double alpha = a.alpha;
    const double alphacold = opt(a.alphacold) ? a.alphacold : alpha / 10.0;
    const double Mx = a.Mx, kerr = a.kerr;
    std::optional<double> orisco = opt(a.risco);
    if (is_ns && !orisco) {
        if (a.nsprop == FB200_NSPROP_DUMMY) {
        } else if (a.nsprop == FB200_NSPROP_NEWT || a.nsprop == FB200_NSPROP_SIBSUN2000) {
            orisco = ss2000_radiusISCO(freqx);
        } else {
            fail(FB200_ERR_INVALID_ARGUMENT, "Wrong nsprop value");
        }
    }
    double rin = opt(a.rin) ? a.rin : r_kerrISCO(Mx, kerr);
    const double rout =
        opt(a.rout) ? a.rout
                    : 0.9 * (rocheLobeVolumeRadiusSemiaxis(Mx / a.Mopt) * semiaxis(Mx + a.Mopt, a.period));
    const double risco = orisco ? *orisco : r_kerrISCO(Mx, kerr);
    if (is_ns && !opt(a.rin)) rin = std::max(Rx, risco);
    auto h_of_r = [Mx](double r) { return std::sqrt(FB_G_GRAV * Mx * r); };

    // ---- DiskStructureArguments, cpp/src/arguments.cpp:30-55 ----
    const double mu = 0.62;  // cpp/include/arguments.hpp:186
    const Oprel oprel = make_oprel(a.opacity, Mx, alpha, mu);
    const double h_in = h_of_r(rin), h_out = h_of_r(rout);
    InitialF ic{};
    if (is_ns && a.initialcond == FB200_IC_QUASISTAT_NS) {
        // cpp/src/ns/ns_arguments.cpp:136-171
        const auto F0 = opt(a.F0), Mdisk0 = opt(a.Mdisk0), Mdot0 = opt(a.Mdot0);
        if (!F0 && !Mdisk0 && !Mdot0) fail(FB200_ERR_RUNTIME, "One of F0, Mdisk0 or Mdot0 must be specified");
        if (Mdot0) {
            const double GMx = Mx * FB_G_GRAV;
            const double Ralfven = a.epsilonAlfven * std::pow(p4(mu_magn) / (GMx * p2(*Mdot0)), 1. / 7.);
            const double h_in_ns = h_of_r(std::max(Ralfven, rin));
            const double coeff = quasistat_coeff(h_in_ns, h_out, oprel);
            ic.kind = FB200_IC_QUASISTAT_NS;
            ic.F0 = *Mdot0 * (h_out - h_in_ns) / h_out * h_in_ns / oprel.f_F(h_in_ns / h_out);
            ic.Mdisk0 = std::pow(ic.F0, 1. - oprel.m) * coeff;
            ic.Mdot0 = *Mdot0;
            ic.h_in_ns = h_in_ns;
        } else if (Mdisk0) {
            fail(FB200_ERR_RUNTIME, "Mdisk0 is not supported for initialcond=quasistat-ns, please let us know if you need it");
        } else {
            fail(FB200_ERR_RUNTIME, "F0 is not supported for initialcond=quasistat-ns, please let us know if you need it");
        }
    } else {
        ic = normalise_initial_F(oprel, a.initialcond, h_in, h_out, opt(a.F0), opt(a.Mdisk0), opt(a.Mdot0), opt(a.powerorder),
                                 opt(a.gaussmu), opt(a.gausssigma));
    }

    // ---- CalculationArguments, cpp/include/arguments.hpp:323-332 ----
    const double tau = opt(a.tau) ? a.tau : a.time / 200;
    const unsigned Nx = a.Nx;
    if (Nx < 4 || Nx > FB200_NX_LIMIT) {
        fail(FB200_ERR_UNSUPPORTED, "Nx must be in [4, " + std::to_string(FB200_NX_LIMIT) + "] for the sm_100a kernels");
    }

    // ---- FreddiState::DiskStructure, cpp/src/freddi_state.cpp:13-53 ----
    d.Nt = static_cast<int32_t>(static_cast<size_t>(std::round(a.time / tau)));
    d.Nx = static_cast<int32_t>(Nx);
    const double GM = FB_G_GRAV * Mx;
    d.GM = GM;
    d.R_g = FB_G_GRAV * Mx / p2(FB_C_LIGHT);
    d.eta = efficiency_of_accretion(kerr);
    d.cosi = std::cos(a.inclination / 180.0 * M_PI);
    // companion star (the geometry itself is built by capi.cu only when the ensemble asks for --starflux)
    d.star_semiaxis = semiaxis(Mx + a.Mopt, a.period);     // cpp/src/freddi_state.cpp:20, cpp/include/orbit.hpp:24-29
    d.star_Topt = a.Topt; d.star_albedo = a.staralbedo;
    d.star_sini = std::sin(a.inclination / 180.0 * M_PI);  // cpp/src/freddi_state.cpp:21, cpp/src/geometry.cpp:117-124
    d.star_cosi = std::cos(a.inclination / 180.0 * M_PI);
    d.star_t0 = a.ephemerist0; d.star_period = a.period;
    d.star_off = 0; d.star_ntri = 0;
    d.distance = a.distance;
    d.cosiOverD2 = d.cosi / p2(a.distance);
    {
        static thread_local GridMemo memo;
        if (!(memo.Nx == Nx && memo.scale == a.gridscale && memo.h_in == h_in && memo.h_out == h_out)) {
            memo.Nx = 0;  // stays invalid if the loop throws
            memo.h.resize(Nx);
            for (size_t i = 0; i < Nx; i++) {
                if (a.gridscale == FB200_GRID_LOG) {
                    memo.h[i] = h_in * pow(h_out / h_in, i / (Nx - 1.));
                } else if (a.gridscale == FB200_GRID_LINEAR) {
                    memo.h[i] = h_in + (h_out - h_in) * i / (Nx - 1.);
                } else {
                    fail(FB200_ERR_INVALID_ARGUMENT, "Wrong gridscale");
                }
            }
            memo.Nx = Nx; memo.scale = a.gridscale; memo.h_in = h_in; memo.h_out = h_out;
        }
        out.h = memo.h;
    }
    std::vector<double> R(Nx);
    for (size_t i = 0; i < Nx; i++) R[i] = p2(out.h[i]) / GM;

    // ---- CurrentState, cpp/src/freddi_state.cpp:56-71 ----
    size_t first = 0;
    initial_F(ic, oprel, out.h, out.F, first);
    d.first0 = static_cast<int32_t>(first);
    d.wind_ctor_Mdot = (first + 1 < Nx) ? (out.F[first + 1] - out.F[first]) / (out.h[first + 1] - out.h[first]) : 0.;  // FreddiState::Mdot_in on the unshifted F
    d.F_in0 = 0.0;
    d.t0 = a.inittime;
    d.Mdot_out = a.Mdotout;

    // ---- irradiation source, cpp/src/freddi_state.cpp:125-133 ----
    if (a.angulardistdisk != FB200_ANGDIST_PLANE && a.angulardistdisk != FB200_ANGDIST_ISOTROPIC)
        fail(FB200_ERR_INVALID_ARGUMENT, "Wrong angular distribution type");
    d.angdist_disk = a.angulardistdisk;

    // ---- wind, cpp/src/freddi_state.cpp:94-122,387-450 (static) ; dynamic winds are evaluated by the kernel ----
    d.windtype = a.windtype;
    for (int k = 0; k < 4; ++k) d.windparams[k] = a.windparams[k];
    const double hf = out.h.front(), hb = out.h.back();
    switch (a.windtype) {
        case FB200_WIND_NO: break;
        case FB200_WIND_SS73C: {
            const double L_edd = 4. * M_PI * FB_M_PROTON * FB_C_LIGHT / FB_THOMSON * GM;
            const double Mdot_crit = L_edd / (p2(FB_C_LIGHT) * d.eta);
            out.wC.assign(Nx, 0.);
            for (size_t i = 0; i < Nx; ++i)
                out.wC[i] = -Mdot_crit / (2 * M_PI * R.front() * R[i]) * (4 * M_PI * p3(out.h[i])) / p2(GM);
            break;
        }
        case FB200_WIND_CAMBIER2013: {
            // The reference reads DiskStructureArguments::Mdot0 here (cpp/src/freddi_state.cpp:408,411), a member its
            // constructor never initialises (cpp/src/arguments.cpp:42-55); the normalised Mdot0 is used instead.
            const double kC = a.windparams[0], R_IC2out = a.windparams[1];
            const double Mdot0 = ic.Mdot0;
            const double m_ch0 = -kC * Mdot0 / (M_PI * p2(R.back()));
            const double L_edd = 4. * M_PI * FB_M_PROTON * FB_C_LIGHT / FB_THOMSON;
            const double Mdot_crit = L_edd / (p2(FB_C_LIGHT) * d.eta);
            const double eta = 0.025 * 33 * Mdot0 / Mdot_crit;
            const double R_iC = R_IC2out * R.back();
            out.wC.assign(Nx, 0.);
            for (size_t i = 0; i < Nx; ++i) {
                const double xi = R[i] / R_iC;
                const double C0 = m_ch0 * (4.0 * M_PI * p3(out.h[i])) / (p2(GM));
                out.wC[i] = C0 * std::pow((1 + p2(0.125 * eta / xi)) / (1 + 1. / (p8(eta) * p2(1 + 262 * p2(xi)))), 1. / 6.) *
                            std::exp(-p2(1 - 1. / std::sqrt(1 + 0.25 / (p2(xi)))) / (2 * xi));
            }
            break;
        }
        case FB200_WIND_TESTA: {
            const double A0 = -a.windparams[0] / p2(hb - hf);
            out.wA.assign(Nx, 0.);
            for (size_t i = 0; i < Nx; ++i) out.wA[i] = A0 * (out.h[i] - hf);
            break;
        }
        case FB200_WIND_TESTB: {
            const double B0 = -a.windparams[0] / p2(hb - hf);
            out.wB.assign(Nx, B0);
            break;
        }
        case FB200_WIND_TESTC: {
            const double C0 = a.windparams[0] * a.Mdotout / (hb - hf);
            const double h_wind_min = hb / 2;
            out.wC.assign(Nx, 0.);
            for (size_t i = 0; i < Nx; ++i) {
                if (out.h[i] > h_wind_min)
                    out.wC[i] = C0 * 0.5 * (std::cos(2. * M_PI * (out.h[i] - h_wind_min) / (hb - h_wind_min)) - 1);
            }
            break;
        }
        case FB200_WIND_SHIELDSOSCIL1986:
        case FB200_WIND_JANIUK2015:
        case FB200_WIND_SHIELDS1986:
        case FB200_WIND_WOODS1996AGN:
        case FB200_WIND_WOODS1996:
        case FB200_WIND_TOY:
            d.wind_dynamic = 1;
            break;
        default:
            fail(FB200_ERR_INVALID_ARGUMENT, "Wrong wind");
    }
    d.wind_static = (!out.wA.empty() || !out.wB.empty() || !out.wC.empty()) ? 1 : 0;
    d.has_AB = (!out.wA.empty() || !out.wB.empty() || a.windtype == FB200_WIND_JANIUK2015) ? 1 : 0;

    // ---- scalars the step and the row read ----
    if (a.boundcond != FB200_BOUNDCOND_TEFF && a.boundcond != FB200_BOUNDCOND_TIRR && a.Thot > 0.)
        fail(FB200_ERR_INVALID_ARGUMENT, "Wrong boundcond");  // thrown by truncateOuterRadius at the first decaying step
    d.opacity = a.opacity;
    d.boundcond = a.boundcond;
    d.tau = tau;
    d.eps = a.eps;
    d.Mx = Mx;
    d.kerr = kerr;
    d.alpha = alpha;
    d.alphacold = alphacold;
    d.mu = mu;
    d.one_minus_m = 1. - oprel.m;
    {
        double c = 1.;
        for (int k = 1; k <= 8; ++k) {  // C(m, k) = C(m, k-1) (m - k + 1) / k
            c *= (oprel.m - (k - 1)) / k;
            d.m_binom[k - 1] = c;
        }
    }
    d.n_op = oprel.n;
    d.D = oprel.D;
    d.Height_coef = oprel.Height_coef;
    d.Height_exp_F = oprel.Height_exp_F;
    d.Height_exp_R = oprel.Height_exp_R;
    d.Thot = a.Thot;
    d.Tirr2Tvishot = std::pow(a.Qirr2Qvishot, 0.25);  // cpp/pywrap/pywrap_freddi_evolution.cpp:148, cpp/src/options.cpp:103
    d.sigma_minus_k = std::pow(alphacold / 0.1, -0.83);
    d.sigma_minus_mx = std::pow(Mx / FB_SOLAR_MASS, -0.40);
    d.sigma_plus_k = std::pow(alpha / 0.1, -0.80);
    d.sigma_plus_mx = std::pow(Mx / FB_SOLAR_MASS, -0.37);
    d.vcf_mx = std::pow(Mx / FB_SOLAR_MASS, -0.012);
    d.Cirr = a.Cirr;
    d.irrindex = a.irrindex;
    d.Cirr_cold = a.Cirrcold;
    d.irrindex_cold = a.irrindexcold;
    d.h2rcold = a.h2rcold;
    d.colourfactor = a.colourfactor;
    d.emin = a.emin;
    d.emax = a.emax;
    // T_GR constants, cpp/src/spectrum.cpp:45-56
    d.tgr_rg = GM / p2(FB_C_LIGHT);
    d.tgr_x0 = std::sqrt(r_kerrISCORg(kerr));
    d.tgr_x1 = 2. * std::cos((std::acos(kerr) - M_PI) / 3.);
    d.tgr_x2 = 2. * std::cos((std::acos(kerr) + M_PI) / 3.);
    d.tgr_x3 = -2. * std::cos(std::acos(kerr) / 3.);

    // ---- FreddiNeutronStarEvolution, cpp/src/ns/ns_evolution.cpp:45-139,322-383 ----
    d.R_dead = 0.;
    d.redshift = 1.;
    d.risco = risco;
    d.angdist_ns = FB200_ANGDIST_ISOTROPIC;
    if (is_ns) {
        // kappa_t (params.at(...) of a missing key is std::out_of_range in the reference; here the array is always filled)
        if (a.kappattype != FB200_KAPPAT_CONST && a.kappattype != FB200_KAPPAT_CORSTEP && a.kappattype != FB200_KAPPAT_ROMANOVA2018)
            fail(FB200_ERR_INVALID_ARGUMENT, "Wrong kappattype");
        d.kappattype = a.kappattype;
        for (int k = 0; k < 4; ++k) d.kappat[k] = a.kappatparams[k];
        d.R_x = Rx;
        if (a.nsgravredshift == FB200_REDSHIFT_OFF) d.redshift = 1.0;
        else if (a.nsgravredshift == FB200_REDSHIFT_ON) d.redshift = 1.0 - 2.0 * d.R_g / Rx;
        else fail(FB200_ERR_INVALID_ARGUMENT, "Wrong nsgravredshift");
        d.R_m_min = std::max(Rx, R[first]);
        d.mu_magn = mu_magn;
        d.freqx = freqx;
        d.R_cor = std::cbrt(GM / p2(2 * M_PI * freqx));
        d.R_dead = a.Rdead > 0. ? a.Rdead : INFINITY;
        d.inverse_beta = a.inversebeta;
        d.epsilon_Alfven = a.epsilonAlfven;
        d.hot_spot_area = a.hotspotarea;
        out.Fmagn.resize(Nx);
        out.dFmagn_dh.resize(Nx);
        out.d2Fmagn_dh2.resize(Nx);
        {
            const double k = d.inverse_beta * p2(mu_magn) / 3.;
            for (size_t i = 0; i < Nx; i++) {
                double brackets;
                if (R[i] < d.R_cor) brackets = -1. + 2. * std::pow(R[i] / d.R_cor, 1.5) - 2. / 3. * p3(R[i] / d.R_cor);
                else brackets = 1. - 2. / 3. * std::pow(d.R_cor / R[i], 1.5);
                out.Fmagn[i] = k / p3(R[i]) * brackets;
            }
        }
        {
            const double k = d.inverse_beta * 2 * p2(mu_magn) * p3(GM);
            for (size_t i = 0; i < Nx; i++) {
                double brackets;
                if (R[i] < d.R_cor) brackets = 1. - std::pow(R[i] / d.R_cor, 1.5);
                else brackets = -1. + std::pow(d.R_cor / R[i], 1.5);
                out.dFmagn_dh[i] = k / p7(out.h[i]) * brackets;
            }
        }
        {
            const double k = d.inverse_beta * 2 * p2(mu_magn) / GM;
            for (size_t i = 0; i < Nx; i++) {
                double brackets;
                if (R[i] < d.R_cor) brackets = -7. + 4. * std::pow(R[i] / d.R_cor, 1.5);
                else brackets = 7. - 10. * std::pow(d.R_cor / R[i], 1.5);
                out.d2Fmagn_dh2[i] = k / p4(R[i]) * brackets;
            }
        }
        if (a.Rdead > 0. && a.Rdead < d.R_cor) fail(FB200_ERR_LOGIC, "R_dead is positive and less than R_cor, it is unacceptably");
        if (a.angulardistns != FB200_ANGDIST_PLANE && a.angulardistns != FB200_ANGDIST_ISOTROPIC)
            fail(FB200_ERR_INVALID_ARGUMENT, "Wrong angular distribution type");
        d.angdist_ns = a.angulardistns;
        // fp, cpp/src/ns/ns_evolution.cpp:342-369
        d.fptype = a.fptype;
        d.fpparams[0] = a.fpparams[0];
        d.fpparams[1] = a.fpparams[1];
        if (a.fptype == FB200_FP_GEOMETRICAL) {
            const double chi = a.fpparams[0] * M_PI / 180.;
            if (chi == 0) d.fptype = FB200_FP_COROTATION_BLOCK;
            d.fpparams[0] = chi;
        } else if (a.fptype < FB200_FP_NO_OUTFLOW || a.fptype > FB200_FP_GEOMETRICAL) {
            fail(FB200_ERR_INVALID_ARGUMENT, "Wrong fptype");
        }
        if (a.nsprop != FB200_NSPROP_DUMMY && a.nsprop != FB200_NSPROP_SIBSUN2000 && a.nsprop != FB200_NSPROP_NEWT)
            fail(FB200_ERR_INVALID_ARGUMENT, "Wrong nsprop");
        d.nsprop = a.nsprop;
        // shift of the initial F by the magnetic torque, cpp/src/ns/ns_evolution.cpp:328-339
        if (d.inverse_beta <= 0.) {
            d.F_in0 = kappa_t(d, R[0]) * p2(mu_magn) / p3(d.R_cor);
            for (size_t i = 0; i < Nx; i++) out.F[i] += d.F_in0;
        } else {
            for (size_t i = 0; i < Nx; i++) out.F[i] += -out.Fmagn[i] + out.Fmagn[0];
        }
    }

    // ---- info ----
    fb200_model_info_t& o = out.info;
    o.GM = GM; o.eta = d.eta; o.R_g = d.R_g; o.cosi = d.cosi; o.distance = d.distance; o.cosiOverD2 = d.cosiOverD2;
    o.tau = tau; o.h_in = out.h.front(); o.h_out = out.h.back(); o.rin = rin; o.rout = rout; o.risco = risco;
    o.F0 = ic.F0; o.Mdisk0 = ic.Mdisk0; o.Mdot0 = ic.Mdot0;
    o.mu_magn = d.mu_magn; o.R_cor = d.R_cor; o.R_dead = d.R_dead; o.inverse_beta = d.inverse_beta; o.R_x = d.R_x;
    o.redshift = d.redshift;
    o.Nx = d.Nx; o.Nt = d.Nt; o.first0 = d.first0;
    return FB200_OK;
}

}  // namespace

int resolve(const fb200_args_t& a, Resolved& out, std::string& err) {
    // `out` may be a reused object (the fill workers keep one per thread): the optional arrays mean "absent" when empty
    out.wA.clear(); out.wB.clear(); out.wC.clear();
    out.Fmagn.clear(); out.dFmagn_dh.clear(); out.d2Fmagn_dh2.clear();
    try {
        return resolve_impl(a, out);
    } catch (const Fail& f) {
        err = f.msg;
        return f.code;
    } catch (const std::exception& e) {
        err = e.what();
        return FB200_ERR_RUNTIME;
    }
}

ModelPlan plan_of(const fb200_args_t& a) {
    ModelPlan p{};
    p.Nx = int(a.Nx);
    p.is_ns = a.is_ns != 0;
    const double tau = std::isnan(a.tau) ? a.time / 200 : a.tau;                 // as resolve_impl (CalculationArguments)
    const double nt = std::round(a.time / tau);
    p.Nt = (nt >= 0. && nt < 2e9) ? int(static_cast<size_t>(nt)) : 0;            // a bad time/tau is reported by resolve()
    p.has_A = a.windtype == FB200_WIND_TESTA;
    p.has_B = a.windtype == FB200_WIND_TESTB;
    p.has_C = a.windtype == FB200_WIND_SS73C || a.windtype == FB200_WIND_CAMBIER2013 || a.windtype == FB200_WIND_TESTC ||
              (p.is_ns && a.inversebeta != 0.);
    p.dynamic = a.windtype == FB200_WIND_SHIELDSOSCIL1986 || a.windtype == FB200_WIND_JANIUK2015 || a.windtype == FB200_WIND_SHIELDS1986 ||
                a.windtype == FB200_WIND_WOODS1996AGN || a.windtype == FB200_WIND_WOODS1996 || a.windtype == FB200_WIND_TOY;
    return p;
}

double star_semiaxis_of(const fb200_args_t& a) { return semiaxis(a.Mx + a.Mopt, a.period); }

int prepare_passband(const double* lam, const double* tr, int n, PassbandTable& out, std::string& err) {
    if (!lam || !tr || n < 2 || n > FB200_MAX_PASSBAND_SAMPLES) {
        err = "a passband needs between 2 and " + std::to_string(FB200_MAX_PASSBAND_SAMPLES) + " tabulated points";
        return FB200_ERR_INVALID_ARGUMENT;
    }
    for (int j = 0; j < n; ++j) {
        if (!(lam[j] > 0.) || !(tr[j] == tr[j])) {
            err = "passband wavelengths must be positive and transmissions not NaN";
            return FB200_ERR_INVALID_ARGUMENT;
        }
    }
    out.a.resize(n);
    out.w.resize(n);
    const double ch_over_kb = FB_C_LIGHT * (FB_H_PLANCK / FB_K_BOLTZ);
    const double double_h_c2 = 2. * FB_H_PLANCK * FB_C_LIGHT * FB_C_LIGHT;
    auto f_dnu = [&](int j) { return tr[j] * FB_C_LIGHT / (lam[j] * lam[j]); };
    // trapz(x, f, 0, n-1), util.cpp:21-30: both ends first, then the interior left to right
    double s = f_dnu(0) * (lam[1] - lam[0]) + f_dnu(n - 1) * (lam[n - 1] - lam[n - 2]);
    for (int j = 1; j <= n - 2; ++j) s += f_dnu(j) * (lam[j + 1] - lam[j - 1]);
    out.t_dnu = 0.5 * s;
    for (int j = 0; j < n; ++j) {
        const double dl = (j == 0) ? lam[1] - lam[0] : (j == n - 1) ? lam[n - 1] - lam[n - 2] : lam[j + 1] - lam[j - 1];
        const double l2 = lam[j] * lam[j];
        out.a[j] = ch_over_kb / lam[j];
        out.w[j] = dl * tr[j] * (double_h_c2 / (l2 * l2 * lam[j]));  // pow<5>(x) = x * (x*x) * (x*x)
    }
    return FB200_OK;
}

}  // namespace fb200
