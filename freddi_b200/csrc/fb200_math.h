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

// ------------------------------------------------------------------------------------------------
// Lean elementary functions for the Planck quadrature (the hottest loop of the row).  They are branch-free
// (so the compiler can interleave the independent stage evaluations of one Cash-Karp step) and are written
// with explicit fma() so that the host build (tests/host_emu) and the device build produce the same bits.
// ------------------------------------------------------------------------------------------------
FB_HD double fb_scale2(double p, int k) {  // p * 2^k for p in [0.5, 2) and a result that stays normal
#ifdef __CUDA_ARCH__
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
#else
    union { double d; long long i; } u;
    u.d = p;
    u.i += (long long)k << 52;
    return u.d;
#endif
}

// 2^(j/32), j = 0..31, correctly rounded.  On the device the table lives in the first 32 doubles of the CTA's
// dynamic shared memory (64 doubles with the tails; every kernel of the library fills it at start, gpu_exec.cuh: stage_cfg): the index
// differs from lane to lane, which shared memory serves and constant memory would serialise.
#define FB_EXP2_TABLE                                                                                   \
    1.0, 1.0218971486541166, 1.0442737824274138, 1.0671404006768237, 1.0905077326652577,               \
    1.1143867425958924, 1.1387886347566916, 1.1637248587775775, 1.189207115002721, 1.215247359980469,  \
    1.241857812073484, 1.2690509571917332, 1.2968395546510096, 1.3252366431597413, 1.3542555469368927, \
    1.383909881963832, 1.4142135623730951, 1.4451808069770467, 1.4768261459394993, 1.5091644275934228, \
    1.5422108254079407, 1.5759808451078865, 1.6104903319492543, 1.645755478153965, 1.681792830507429,  \
    1.718619298122478, 1.7562521603732995, 1.7947090750031072, 1.8340080864093424, 1.8741676341103,    \
    1.9152065613971474, 1.9571441241754002
// 2^(j/32) - (double)2^(j/32): the tail that brings the table product back to ~0.5 ulp (used by fb_pow_pos)
#define FB_EXP2_TABLE_LO \
    0.0, 5.109225028973444e-17, 8.551889705537965e-17, -7.899853966841582e-17, \
    -3.046782079812471e-17, 1.0410278456845571e-16, 8.912812676025408e-17, 3.8292048369240935e-17, \
    3.982015231465646e-17, -7.712630692681488e-17, 4.658027591836937e-17, 2.667932131342186e-18, \
    2.5382502794888315e-17, -2.8587312100388614e-17, 7.70094837980299e-17, -6.770511658794786e-17, \
    -9.667293313452913e-17, -3.0237581349939873e-17, -3.483994556892796e-17, -1.016455327754295e-16, \
    7.949834809697621e-17, -1.0136916471278304e-17, 2.4707192569797888e-17, -1.0125679913674773e-16, \
    8.199010020581497e-17, -1.851380418263111e-17, 2.960140695448873e-17, 1.8227458427912087e-17, \
    3.283107224245627e-17, -6.122763413004143e-17, -1.0619946056195963e-16, 8.960767791036668e-17
#ifdef __CUDACC__
extern __shared__ __align__(16) double fb_smem[];
#endif
static const double fb_exp2_table_host[64] = {FB_EXP2_TABLE, FB_EXP2_TABLE_LO};
#ifdef __CUDA_ARCH__
#define FB_EXP2_TAB(j) fb_smem[j]
#else
#define FB_EXP2_TAB(j) fb_exp2_table_host[j]
#endif

// exp(x) for -700 <= x <= 709, within ~1 ulp: k = rint(x log2 e), r = x - k ln2 (two-piece ln2), 13th-degree
// Taylor polynomial on |r| <= 0.347 (truncation 4e-18), scaling through the exponent field.
FB_HD double fb_exp(double x) {
    const double shift = 6755399441055744.0;  // 1.5 * 2^52: adding it rounds to the nearest integer
    const double t = fma(x, 1.4426950408889634074, shift);
    const double k = t - shift;
#ifdef __CUDA_ARCH__
    const int ki = __double2loint(t);
#else
    union { double d; long long i; } u;
    u.d = t;
    const int ki = (int)(u.i & 0xffffffffLL);
#endif
    double r = fma(k, -6.93147180369123816490e-01, x);
    r = fma(k, -1.90821492927058770002e-10, r);
    double p = 1.6059043836821613e-10;           // 1/13!
    p = fma(p, r, 2.08767569878681e-09);         // 1/12!
    p = fma(p, r, 2.505210838544172e-08);        // 1/11!
    p = fma(p, r, 2.755731922398589e-07);        // 1/10!
    p = fma(p, r, 2.7557319223985893e-06);       // 1/9!
    p = fma(p, r, 2.48015873015873e-05);         // 1/8!
    p = fma(p, r, 1.984126984126984e-04);        // 1/7!
    p = fma(p, r, 1.388888888888889e-03);        // 1/6!
    p = fma(p, r, 8.333333333333333e-03);        // 1/5!
    p = fma(p, r, 4.1666666666666664e-02);       // 1/4!
    p = fma(p, r, 1.6666666666666666e-01);       // 1/3!
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return fb_scale2(p, ki);
}

// Full-precision FP64 literals cost two UMOV / IMAD.MOV each on sm_100a every time they are used (DFMA takes no 64-bit immediate
// and ptxas rematerialises them); from __constant__ memory one LDCU.128 fetches two.  FB_KC(name) is the table entry on the
// device and the literal on the host — one list, so the two cannot drift apart.
#define FB_KC_LIST(X)                                                                                                     \
    X(EXP_32_LN2, 46.16624130844683) X(EXP_LN2_32_HI, -0.021660849392446835)                                             \
    X(EXP_LN2_32_LO, -5.145609244655338e-14) X(EXP_Q6, 1.388888888888889e-03)                                            \
    X(EXP_Q5, 8.333333333333333e-03) X(EXP_Q4, 4.1666666666666664e-02)                                                   \
    X(EXP_Q3, 1.6666666666666666e-01) X(CK_B1, 37. / 378.)                                                               \
    X(CK_B3, 250. / 621.) X(CK_B4, 125. / 594.)                                                                          \
    X(CK_B6, 512. / 1771.) X(CK_DB1, 37. / 378. - 2825. / 27648.)                                                        \
    X(CK_DB3, 250. / 621. - 18575. / 48384.) X(CK_DB4, 125. / 594. - 13525. / 55296.)                                    \
    X(CK_DB5, -277. / 14336.) X(CK_DB6, 512. / 1771. - 1. / 4.)                                                          \
    X(CK_C3, 3. / 10.) X(CK_C4, 3. / 5.)                                                                                 \
    X(CK_FLOOR, 0.00032) X(CK_NINE_TENTHS, 9. / 10.)                                                                     \
    X(CK_M_THIRD, -1. / 3.) X(CK_M_FIFTH, -1. / 5.)                                                                      \
    X(CK_FIFTH, 1. / 5.) X(CK_HUGE, 1e300)                                                                               \
    X(ATH_21, 0.047619047619047616) X(ATH_19, 0.05263157894736842)                                                       \
    X(ATH_17, 0.058823529411764705) X(ATH_15, 0.06666666666666667)                                                       \
    X(ATH_13, 0.07692307692307693) X(ATH_11, 0.09090909090909091)                                                        \
    X(ATH_9, 0.1111111111111111) X(ATH_7, 0.14285714285714285)                                                           \
    X(ATH_5, 0.2) X(ATH_3, 0.3333333333333333)                                                                           \
    X(TWO_OVER_LN2_HI, 2.8853900817779268) X(TWO_OVER_LN2_LO, 4.0710547481862066e-17)                                    \
    X(LN2_HI, 0.6931471805599453) X(LN2_LO, 2.3190468138462996e-17)                                                      \
    X(INVFACT_11, 2.505210838544172e-08) X(INVFACT_10, 2.755731922398589e-07)                                            \
    X(INVFACT_9, 2.7557319223985893e-06) X(INVFACT_8, 2.48015873015873e-05)                                              \
    X(INVFACT_7, 1.984126984126984e-04) X(CK_TOL, 1e-4)
enum {
#define FB_KC_ENUM(n, v) FBK_##n,
    FB_KC_LIST(FB_KC_ENUM)
#undef FB_KC_ENUM
    FBK_COUNT
};
#ifdef __CUDACC__
static __constant__ __align__(16) double fb_kc[FBK_COUNT] = {
#define FB_KC_VAL(n, v) v,
    FB_KC_LIST(FB_KC_VAL)
#undef FB_KC_VAL
};
#endif
#define FB_KC_HOST(n, v) static const double FBKV_##n = v;
FB_KC_LIST(FB_KC_HOST)
#undef FB_KC_HOST
#ifdef __CUDA_ARCH__
#define FB_KC(n) fb_kc[FBK_##n]
#else
#define FB_KC(n) FBKV_##n
#endif
// FB_KL(name): the literal on both sides — for the places where the table entry measured slower than the immediates (the step-size
// power of the quadrature controller and the controller's own constants: 3.82 -> 3.80 M model-steps/s)
#define FB_KL(n) FBKV_##n

// exp(x) for -700 <= x <= 709 for the Planck function, ~1 ulp with 11 FP64 operations: n = rint(32 x / ln2),
// exp(x) = 2^(n >> 5) * 2^((n & 31)/32) * exp(r), |r| <= ln2/64, expm1(r) by a 6th-degree polynomial (truncation 4e-18).
FB_HD double fb_exp_tab(double x) {
    const double shift = 6755399441055744.0;  // 1.5 * 2^52
    const double t = fma(x, FB_KC(EXP_32_LN2), shift);  // 32 / ln2
    const double n = t - shift;
#ifdef __CUDA_ARCH__
    const int ni = __double2loint(t);
#else
    union { double d; long long i; } u;
    u.d = t;
    const int ni = (int)(u.i & 0xffffffffLL);
#endif
    double r = fma(n, FB_KC(EXP_LN2_32_HI), x);      // ln2/32, leading 37 bits (n * hi is exact)
    r = fma(n, FB_KC(EXP_LN2_32_LO), r);             // ln2/32, tail
    double q = fma(r, FB_KC(EXP_Q6), FB_KC(EXP_Q5));
    q = fma(q, r, FB_KC(EXP_Q4));
    q = fma(q, r, FB_KC(EXP_Q3));
    q = fma(q, r, 0.5);
    const double p = fma(q * r, r, r);               // expm1(r)
    const double tj = FB_EXP2_TAB(ni & 31);
    return fb_scale2(fma(tj, p, tj), ni >> 5);
}

// 1/y for normal positive y, ~1 ulp (device: hardware seed + two Newton steps; host: IEEE division)
FB_HD double fb_rcp(double y) {
#ifdef __CUDA_ARCH__
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y));
    double e = fma(-y, r, 1.0);
    r = fma(r, e, r);
    e = fma(-y, r, 1.0);
    r = fma(r, e, r);
    return r;
#else
    return 1.0 / y;
#endif
}

// 1/y to ~1e-14 (hardware seed 2^-23 + ONE Newton step): for quotients that only feed a comparison or a correction term that is
// itself small (the controller's error ratio; x = (y_old - y) / y of the incremental K update, |x| < 2^-6)
FB_HD double fb_rcp14(double y) {
#ifdef __CUDA_ARCH__
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y));
    const double e = fma(-y, r, 1.0);
    return fma(r, e, r);
#else
    return 1.0 / y;
#endif
}

// x^a for normal positive x and |a log2 x| < 1000, ~1 ulp — the opacity closure W ~ |F|^(1-m) and the disc height
// ~ F^(3/20), evaluated once per ring per Picard iteration.  Same structure as a libm pow without its special-case
// branches: log2 x in double-double (x = 2^e m, m in [sqrt(1/2), sqrt 2), log m = 2 atanh((m-1)/(m+1)) with the leading
// term carried as hi+lo and the s^3.. tail in double), y = a log2 x in double-double, 2^y = 2^k exp(f ln2).
// Anything else (zero, negative, NaN, inf, subnormal) goes to the library pow.
FB_HD double fb_pow_pos(double x, double a) {
#ifdef __CUDA_ARCH__
    int hi = __double2hiint(x);
    const int lo = __double2loint(x);
#else
    union { double d; long long i; } u_;
    u_.d = x;
    int hi = (int)(u_.i >> 32);
    const int lo = (int)(u_.i & 0xffffffffLL);
#endif
    if ((unsigned)(hi - 0x00100000) >= 0x7fe00000u) return pow(x, a);  // not a positive normal number
    int e = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;
    if (hi >= 0x3ff6a09e) { hi -= 0x00100000; e += 1; }
#ifdef __CUDA_ARCH__
    const double m = __hiloint2double(hi, lo);
#else
    u_.i = ((long long)hi << 32) | (unsigned int)lo;
    const double m = u_.d;
#endif
    // s = (m - 1) / (m + 1) as s_hi + s_lo
    const double um = m - 1.0;                  // exact
    const double v_hi = m + 1.0;
    const double v_lo = m - (v_hi - 1.0);       // exact residual of the sum
    const double r = fb_rcp(v_hi);
    const double s_hi = um * r;
    double t = fma(-s_hi, v_hi, um);
    t = fma(-s_hi, v_lo, t);
    const double s_lo = t * r;
    // atanh(s) = s + s (s^2/3 + s^4/5 + ... + s^20/21), |s| <= 0.1716
    const double s2 = s_hi * s_hi;
    double q = FB_KC(ATH_21);
    q = fma(q, s2, FB_KC(ATH_19));
    q = fma(q, s2, FB_KC(ATH_17));
    q = fma(q, s2, FB_KC(ATH_15));
    q = fma(q, s2, FB_KC(ATH_13));
    q = fma(q, s2, FB_KC(ATH_11));
    q = fma(q, s2, FB_KC(ATH_9));
    q = fma(q, s2, FB_KC(ATH_7));
    q = fma(q, s2, FB_KC(ATH_5));
    q = fma(q, s2, FB_KC(ATH_3));
    const double a_lo = fma(s_hi * s2, q, s_lo);        // tail + s_lo
    // log2 m = (2/ln2) (s_hi + a_lo), double-double product
    const double c_hi = FB_KC(TWO_OVER_LN2_HI), c_lo = FB_KC(TWO_OVER_LN2_LO);
    const double p_hi = s_hi * c_hi;
    const double p_lo = fma(s_hi, c_hi, -p_hi) + fma(s_hi, c_lo, a_lo * c_hi);
    // log2 x = e + p_hi + p_lo
    const double ed = (double)e;
    const double l_hi = ed + p_hi;
    const double l_lo = (p_hi - (l_hi - ed)) + p_lo;
    // y = a log2 x
    const double y_hi = a * l_hi;
    const double y_lo = fma(a, l_hi, -y_hi) + a * l_lo;
    // 2^y = 2^(n >> 5) * 2^((n & 31)/32) * exp(f ln2), n = rint(32 y), |f| <= 1/64
    const double shift = 6755399441055744.0;
    const double tk = fma(y_hi, 32.0, shift);
    const double nk = tk - shift;
#ifdef __CUDA_ARCH__
    const int ni = __double2loint(tk);
#else
    u_.d = tk;
    const int ni = (int)(u_.i & 0xffffffffLL);
#endif
    if (ni > 32000 || ni < -32000) return pow(x, a);
    const double f = fma(nk, -0.03125, y_hi) + y_lo;   // exact subtraction, then the low part
    const double g = fma(f, FB_KC(LN2_HI), f * FB_KC(LN2_LO));
    double w = fma(g, FB_KC(EXP_Q6), FB_KC(EXP_Q5));
    w = fma(w, g, FB_KC(EXP_Q4));
    w = fma(w, g, FB_KC(EXP_Q3));
    w = fma(w, g, 0.5);
    const double pm1 = fma(w * g, g, g);               // expm1(g)
    const double tj = FB_EXP2_TAB(ni & 31);
    return fb_scale2(tj + fma(tj, pm1, FB_EXP2_TAB(32 + (ni & 31))), ni >> 5);  // small terms first: one final rounding
}

// base^p for base in [1e-5, 1e300], |p| <= 1/3, relative accuracy ~1e-12: only steers the step size of the
// adaptive quadrature, where it needs no more.
FB_HD double fb_pow_coarse(double base, double p) {
#ifdef __CUDA_ARCH__
    int hi = __double2hiint(base);
    const int lo = __double2loint(base);
#else
    union { double d; long long i; } u;
    u.d = base;
    int hi = (int)(u.i >> 32);
    const int lo = (int)(u.i & 0xffffffffLL);
#endif
    int e = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;       // mantissa in [1, 2)
    if (hi >= 0x3ff6a09e) { hi -= 0x00100000; e += 1; }  // fold [sqrt2, 2) to [sqrt(1/2), 1)
#ifdef __CUDA_ARCH__
    const double m = __hiloint2double(hi, lo);
#else
    u.i = ((long long)hi << 32) | (unsigned int)lo;
    const double m = u.d;
#endif
    // log2(m) = (2/ln2) atanh(s), s = (m-1)/(m+1), |s| <= 0.1716
    const double s = (m - 1.0) * fb_rcp(m + 1.0);
    const double s2 = s * s;
    double q = FB_KL(ATH_15);
    q = fma(q, s2, FB_KL(ATH_13));
    q = fma(q, s2, FB_KL(ATH_11));
    q = fma(q, s2, FB_KL(ATH_9));
    q = fma(q, s2, FB_KL(ATH_7));
    q = fma(q, s2, FB_KL(ATH_5));
    q = fma(q, s2, FB_KL(ATH_3));
    q = fma(q * s2, s, s);
    const double lg = fma(q, FB_KL(TWO_OVER_LN2_HI), (double)e);  // 2/ln2
    // 2^(p lg)
    const double y = p * lg;
    const double shift = 6755399441055744.0;
    const double t = y + shift;
    const double k = t - shift;
#ifdef __CUDA_ARCH__
    const int ki = __double2loint(t);
#else
    u.d = t;
    const int ki = (int)(u.i & 0xffffffffLL);
#endif
    const double f = (y - k) * FB_KL(LN2_HI);  // ln2
    double w = FB_KL(INVFACT_11);            // 1/11!
    w = fma(w, f, FB_KL(INVFACT_10));
    w = fma(w, f, FB_KL(INVFACT_9));
    w = fma(w, f, FB_KL(INVFACT_8));
    w = fma(w, f, FB_KL(INVFACT_7));
    w = fma(w, f, FB_KL(EXP_Q6));
    w = fma(w, f, FB_KL(EXP_Q5));
    w = fma(w, f, FB_KL(EXP_Q4));
    w = fma(w, f, FB_KL(EXP_Q3));
    w = fma(w, f, 0.5);
    w = fma(w, f, 1.0);
    w = fma(w, f, 1.0);
    return fb_scale2(w, ki);
}

// err^(-1/3) (third) or err^(-1/5) for err in [3.2e-4, 100]: the step-size factors of the quadrature controller
// (odeint's default step adjuster).  An FP32 seed from the special-function unit (2^(p log2 err), ~1e-6) and two
// Newton steps for y^-n = err in FP64 (y <- y (n + 1 - err y^n) / n, error 3 e^2 per step: 1e-6 -> 3e-12 -> 3e-23): 14 FP64
// instructions, accurate to rounding, against the 35 of fb_pow_coarse — and the seed runs beside the FP64 pipe.
FB_HD double fb_pow_step_factor(double err, bool third) {
    const float p = third ? -0.33333334f : -0.2f;
#ifdef __CUDA_ARCH__
    double y = (double)exp2f(p * __log2f((float)err));
#else
    double y = (double)exp2f(p * log2f((float)err));
#endif
    const double c1 = third ? 4. : 6., c2 = third ? 0.3333333333333333 : 0.2;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const double y2 = y * y, y3 = y2 * y;
        const double yn = third ? y3 : y3 * y2;
        y = (y * fma(-err, yn, c1)) * c2;
    }
    return y;
}

FB_HD double angular_dist(int type, double mu) { return type == FB200_ANGDIST_ISOTROPIC ? 1.0 : 2.0 * mu; }

}  // namespace fb200
#endif
