// Small host+device helpers shared by resolve.cpp and the kernels.
#ifndef FB200_MATH_H
#define FB200_MATH_H

#include <math.h>

#include "../../include/freddi_b200.h"
#include "fb200_model.h"

#ifdef __CUDACC__
#define FB_HD __host__ __device__ __forceinline__
#else
#define FB_HD inline
#endif

namespace fb200 {

// Integer powers in the multiplication order of boost::math::pow<N> (the reference's m::pow<N>,
// cpp/include/util.hpp:11-14): even N: p=pow<N/2>, p*p; odd N: x*p*p.
FB_HD double p2(double x) { return x * x; }
FB_HD double p3(double x) { return x * x * x; }
FB_HD double p4(double x) { const double p = x * x; return p * p; }
FB_HD double p5(double x) { const double p = x * x; return x * p * p; }
FB_HD double p6(double x) { const double p = x * x * x; return p * p; }
FB_HD double p7(double x) { const double p = x * x * x; return x * p * p; }
FB_HD double p8(double x) { double p = x * x; p = p * p; return p * p; }

// kappa_t(R): cpp/src/ns/ns_evolution.cpp:7-42
FB_HD double kappa_t(const fb200_model_dev_t& m, double R) {
    switch (m.kappattype) {
        case FB200_KAPPAT_CONST: return m.kappat[0];
        case FB200_KAPPAT_CORSTEP: return (R > m.R_cor) ? m.kappat[1] : m.kappat[0];
        default: {
            const double R_to_Rcor = R / m.R_cor;
            if (R_to_Rcor > 1.0) {
                const double fastness = pow(R_to_Rcor, 1.5);
                const double feff = m.kappat[2] * pow(fastness, m.kappat[3]);
                const double out_minus_wind = m.kappat[1] - feff * pow(m.epsilon_Alfven, 3.5);
                return (out_minus_wind < 0) ? 0.0 : out_minus_wind;
            }
            return m.kappat[0];
        }
    }
}

// fp(R): fraction of the accretion rate reaching the star, cpp/src/ns/ns_evolution.cpp:144-225
FB_HD double ns_fp(const fb200_model_dev_t& m, double R) {
    if (R <= m.R_x || R <= m.risco) return 1.0;
    const double x = R / m.R_cor;
    switch (m.fptype) {
        case FB200_FP_NO_OUTFLOW: return 1.;
        case FB200_FP_PROPELLER: return 0.;
        case FB200_FP_COROTATION_BLOCK: return (x > 1.) ? 0. : 1.;
        case FB200_FP_EKSI_KUTLU2010: {
            const double fastness2 = p3(x);
            double p = (1. - 1. / fastness2);
            if (p < 0) p = 0;
            return 1. - 1.5 * sqrt(p) + 0.5 * pow(p, 1.5);
        }
        case FB200_FP_ROMANOVA2018: {
            const double fastness = pow(x, 1.5);
            double f = 0;
            if (fastness >= 1) f = m.fpparams[0] * pow(fastness, m.fpparams[1]);
            if (f > 1) f = 1.;
            return 1. - f;
        }
        default: {  // geometrical with chi != 0 (chi == 0 is mapped to corotation-block by the resolver)
            const double chi = m.fpparams[0];
            const double mdot_factor = pow(x, -3.5);
            if (mdot_factor > 1.) return 1;
            const double tmp = mdot_factor - p2(cos(chi));
            if (tmp < 0.) return 0;
            return 1. - 2. / FB_PI * acos(sqrt(tmp) / sin(chi));
        }
    }
}

// eta_ns(Rm): cpp/src/ns/ns_evolution.cpp:230-319; R_in = R[first]
FB_HD double ns_eta(const fb200_model_dev_t& m, double Rm, double R_in) {
    const double Rx = m.R_x, Risco = m.risco;
    // 0: Rx furthest, 1: Risco furthest, 2: Rm furthest (also when Rm is NaN)
    int which = 2;
    if ((Rx >= Risco) && (Rx >= Rm)) which = 0;
    else if ((Risco >= Rx) && (Risco >= Rm)) which = 1;
    const double Rsch = 2.0 * m.R_g;
    const double omega_ns = m.freqx * 2 * FB_PI;
    if (m.nsprop == FB200_NSPROP_DUMMY) {
        return m.R_g * (1. / Rx - 0.5 / R_in);
    }
    if (m.nsprop == FB200_NSPROP_NEWT) {
        if (which == 2) {
            double Reff = Rm;
            if (Rm > m.R_cor) Reff = m.R_cor;
            const double omega_Kepl_Rin = sqrt(m.GM / p3(Reff));
            return Rsch / 2.0 / Rx * (1.0 - Rx / Reff) +
                   p2(omega_ns) / 2.0 / p2(FB_C_LIGHT) * (Rx - Reff) * (Rx + Reff) +
                   Rsch / 4.0 / Reff * p2(1.0 - omega_ns / omega_Kepl_Rin);
        }
        const double omega_Kepl_Rx = sqrt(m.GM / p3(Rx));
        return Rsch / 4.0 / Rx * p2(1.0 - omega_ns / omega_Kepl_Rx);
    }
    // sibgatullinsunyaev2000
    const double freqx_kHz = m.freqx / 1000.0;
    const double eta_ns_plus_disk = 0.213 - 0.153 * freqx_kHz + 0.02 * p2(freqx_kHz);
    const double ns_to_ns_plus_disk = 0.737 - 0.312 * freqx_kHz - 0.19 * p2(freqx_kHz);
    const double etaSS2000 = eta_ns_plus_disk * ns_to_ns_plus_disk;
    if (which != 2) return etaSS2000;
    const double omega_Kepl_Rx = sqrt(m.GM / p3(Rx));
    const double k = etaSS2000 / p2(1.0 - omega_ns / omega_Kepl_Rx);
    double Reff = Rm;
    if (Rm > m.R_cor) Reff = m.R_cor;
    const double omega_Kepl_Rin = sqrt(m.GM / p3(Reff));
    return sqrt(1.0 - Rsch / Reff) - sqrt(1.0 - Rsch / Rx) +
           p2(omega_ns) / 2.0 / p2(FB_C_LIGHT) * (Rx - Reff) * (Rx + Reff) + k * p2(1.0 - omega_ns / omega_Kepl_Rin);
}

FB_HD double angular_dist(int type, double mu) { return type == FB200_ANGDIST_ISOTROPIC ? 1.0 : 2.0 * mu; }

}  // namespace fb200
#endif
