// The disc-evolution engine: one implicit time step + one light-curve row for ONE model whose radial
// points are spread over the threads of a CTA (one point per thread).
//
// The code is a template over an execution policy `Ex`:
//   * evolve.cu instantiates it with the CUDA policy (threads of a CTA, __syncthreads, warp-shuffle
//     reductions, shared memory) — that is the product;
//   * tests/host_emu instantiates it with a sequential policy so the `-m "not gpu"` tests can run the
//     very same step logic on a CPU against the oracle (test infrastructure only, never shipped).
// Orchestration code below is "uniform": on the GPU every thread of the CTA executes it with identical
// values; per-point work is in `ex.points([&](int i, int k) {...})` lambdas, and data written by
// one points() phase is only read by another thread after `ex.sync()`.
//
// Reference restated (paths relative to the reference repository):
//   step                  FreddiEvolution::step           cpp/src/freddi_evolution.cpp:17-29
//                         FreddiState::step               cpp/src/freddi_state.cpp:147-153
//   solve_step            nonlinear_diffusion_nonuniform_wind_1_2   cpp/src/nonlinear_diffusion.cpp:31-89
//   wunc                  FreddiEvolution::wunction       cpp/src/freddi_evolution.cpp:77-83
//   truncate_outer        FreddiEvolution::truncateOuterRadius      cpp/src/freddi_evolution.cpp:32-74
//   truncate_inner        FreddiNeutronStarEvolution::truncateInnerRadius  cpp/src/ns/ns_evolution.cpp:431-468
//   structure/row         FreddiState lazy getters        cpp/src/freddi_state.cpp:156-312, cpp/include/freddi_state.hpp:377-424
//                         row columns                     cpp/src/output.cpp:197-233, cpp/src/ns/ns_output.cpp:6-19
//   planck_*              Spectrum::*                     cpp/src/spectrum.cpp:10-68
//   dynamic winds         *Wind::update                   cpp/src/freddi_state.cpp:457-654
#ifndef FB200_DISC_ENGINE_H
#define FB200_DISC_ENGINE_H

#include "fb200_math.h"
#include "lx_table.h"

#ifndef __CUDACC__
// host-only builds (tests/host_emu): the two CUDA vector helpers the engine uses
struct double2 { double x, y; };
inline double2 make_double2(double x, double y) { return double2{x, y}; }
#endif

namespace fb200 {

#define FB200_LX_PRUNE 100.0 /* see structure_and_row: upper limit of the per-row pruning exponent of the Lx quadrature */
#define FB200_MAXQ 44 /* reduction slots: Mdisk, Lx, Fnu[16], Fnu_cold[16], passbands[4], passbands_cold[4], Lx trial count */

// Ensemble-wide output configuration (device copy of fb200_config_t's output part)
struct OutCfg {
    int n_lambdas;
    int cold_disk;
    int compute_lx;
    int record_F;
    int n_cols;
    int n_rows;  // rows allocated per model
    int lx_prune;  // 1: skip the Lx quadrature of rings that cannot contribute (bound in structure_and_row)
    int n_passbands;
    double lambdas[FB200_MAX_LAMBDAS];
    // tabulated passbands (resolve.hpp PassbandTable), samples of passband p at [pb_off[p], pb_off[p+1]) of pb_a / pb_w
    int pb_off[FB200_MAX_PASSBANDS + 2];
    double pb_tdnu[FB200_MAX_PASSBANDS];
    const double* pb_a;
    const double* pb_w;
    // companion star (--starflux): triangle tables of all geometries back to back (star.hpp StarGeometry::packed)
    const double* star_data;
    int star_flux;
    int _pad;
    // the band integral of the Planck function as a table in x1 = h nu1 / kT (lx_table.h), or nullptr: walk every ring.
    // Valid for the models whose fb200_model_dev_t::lx_tab is set (their emax / emin is the table's band ratio).
    const double* lx_table;
};

// Mutable per-model state carried between evolve calls (FreddiState::CurrentState, cpp/include/freddi_state.hpp:242-258)
struct ModelState {
    double t, F_in, Mdot_in_prev;
    int first, last, i_t, status, rows_valid, row0_done;
    int target_i_t;  // i_t this launch runs to (set by the first work item of the launch)
    int k_valid;     // EnsembleDev::K holds the closure K of the state in F (DiscEngine::k_valid carried between work items)
    long long picard_iters;
    long long lx_trials;  // try_step calls of the Lx quadrature, summed over points and rows (roofline bookkeeping)
    long long ring_steps; // sum over the solves of (last - first): the unknowns the implicit solves really had (roofline bookkeeping)
    long long ring_iters; // sum over the solves of (last - first) x Picard iterations
    long long row_rings;  // sum over the rows written of the hot-zone rings (last - first + 1) their columns were evaluated on
    long long pow_evals;  // closure pow() calls the Picard loops really made (set-up, first iterations with a source term, series fall-backs)
    // What the last BasicWind::update() of a dynamic wind saw (cpp/src/freddi_state.cpp:147-153,457-654): the reference's
    // windB()/windC() getters and Mdot_wind() return the arrays that update left, i.e. values of the state BEFORE the solve.
    double wind_Mdot;
    int wind_first, wind_last;
};

// One per (model, row) of an ensemble with dynamic winds: the arguments of the update() whose values the reference's wind
// getters hold at that row.  Row 0 holds the constructor's update (BH Mdot_in formula on the unshifted initial F, T12 of
// SURVEY.md 9).  update() rewrites the rings of [first, last] only, so a ring outside keeps the value of the last update
// that covered it: the read-out kernels walk the records backwards (fields.cu).
struct WindRecord {
    double Mdot;
    int first, last;
};

// Bookkeeping the engine reports through its execution policy (ex.count / ex.note_*) instead of carrying it in the
// per-thread copy of ModelState: on the GPU that copy lives in registers for the whole step loop, and 14 registers of
// counters cost spills in the Lx walk.  The policy adds them to ModelState when the work item ends.
enum { FB_CNT_PICARD = 0, FB_CNT_LX_TRIALS, FB_CNT_RING_STEPS, FB_CNT_RING_ITERS, FB_CNT_POW_EVALS, FB_CNT_POW_FALLBACKS, FB_CNT_ROW_RINGS, FB_CNT_N };


// Per-point arrays of one model (length FB200_NX_MAX).  The engine only needs `S.name[i]`; each execution
// policy supplies its own `Ex::Arrays` type with these members.  PtrArrays is the plain-pointer version (host
// policy).  On the GPU the arrays touched in every Picard iteration are shared-memory windows at compile-time
// offsets (so that accesses compile to LDS/STS with immediate offsets and cost no registers), the ones read
// once per step (grid, row constants, interior coefficients) live in an L2-resident global scratch slot.
struct PtrArrays {
    const double* h;                   // grid (global: the ensemble's h array)
    double *R;                         // radius (scratch)
    double *F;                         // state (shared)
    double *hn;                        // h^n / (1-m) / D of the opacity closure (shared)
    double *ca, *cb, *cc, *frac;       // interior finite-difference coefficients, nonlinear_diffusion.cpp:46-51 (scratch)
    double *c1, *hm175, *hotR, *Gb;    // per-point constants of the row, see init_points (scratch)
    double *sgA, *qB, *tv4;            // grid-only factors of the row's structure pass: Sigma / W, (Qx / sigma_SB) / (Kirr L), (Tvis / F^(1/4))^4
    // this step's tridiagonal system  l_i y_{i-1} + d_i y_i + u_i y_{i+1} = rhs_i  (l = -a_i, u = -b_i, d = c0_i + K_i);
    // s_l, s_u, s_d, s_rhs and the block-elimination results p_a, p_c, p_g use the swizzled index pidx() (shared)
    double *s_l, *s_u, *s_d, *s_rhs, *s_cc, *s_frac;
    double *s_K;                       // K_i = frac_i W_i / y_i of the Picard iteration (shared, plain index)
    double *p_a, *p_c, *p_g;
    double2 *lu[2];                    // PCR ping-pong of the reduced system: (lower, upper), diagonal normalised to 1
    double *rr[2];                     // PCR ping-pong of the reduced system: right-hand side
    // row temporaries alias the PCR buffers (never live at the same time)
    double *Tph, *Tirr, *Sigma, *Tvis, *TphX, *aux, *wq;
};

// ------------------------------------------------------------------------------------------------
// Spectrum (cpp/src/spectrum.cpp:10-37, constants cpp/include/spectrum.hpp:15-18)
// ------------------------------------------------------------------------------------------------
#define FB_DOUBLE_H_OVER_C2 (2. * FB_H_PLANCK / (FB_C_LIGHT * FB_C_LIGHT))
#define FB_H_OVER_KB (FB_H_PLANCK / FB_K_BOLTZ)
#define FB_DOUBLE_H_C2 (2. * FB_H_PLANCK * FB_C_LIGHT * FB_C_LIGHT)
#define FB_CH_OVER_KB (FB_C_LIGHT * (FB_H_PLANCK / FB_K_BOLTZ))

// B_lambda with the lambda-only factor pre-divided: pref = double_h_c2 / lambda^5 (same operation order
// as `double_h_c2 / pow<5>(lambda) / (exp(...) - 1)`)
FB_HD double planck_lambda_pref(double T, double lambda, double pref) {
    if (!(T > 0.)) return 0.;
    const double a = FB_CH_OVER_KB * fb_rcp(lambda * T);
    const double b = pref * fb_rcp(fb_exp_tab((a < 700.) ? a : 700.) - 1.);
    return (a < 700.) ? b : 0.;  // beyond exp(700) the value is < 1e-300 of the Rayleigh-Jeans level
}

// sum_j w_j / (exp(a_j / T) - 1) over the samples of one tabulated passband: the lambda-trapezoid of
// B_lambda(T, lambda_j) * transmission_j (cpp/include/freddi_state.hpp:408-417) for ONE ring, so that the radial and
// the wavelength trapezoids can be exchanged (both are plain weighted sums; all terms are >= 0) and a ring's
// temperature stays in a register for the whole passband.  0 when !(T > 0), as Planck_lambda (spectrum.cpp:17-22).
FB_HD double passband_ring_sum(double T, const double* __restrict__ a, const double* __restrict__ w, int n) {
    if (!(T > 0.)) return 0.;
    const double invT = fb_rcp(T);
    double s0 = 0., s1 = 0.;
    int j = 0;
    for (; j + 1 < n; j += 2) {  // two independent exp/reciprocal chains in flight
        const double x0 = a[j] * invT, x1 = a[j + 1] * invT;
        const double b0 = w[j] * fb_rcp(fb_exp_tab((x0 < 700.) ? x0 : 700.) - 1.);
        const double b1 = w[j + 1] * fb_rcp(fb_exp_tab((x1 < 700.) ? x1 : 700.) - 1.);
        s0 += (x0 < 700.) ? b0 : 0.;
        s1 += (x1 < 700.) ? b1 : 0.;
    }
    if (j < n) {
        const double x0 = a[j] * invT;
        const double b0 = w[j] * fb_rcp(fb_exp_tab((x0 < 700.) ? x0 : 700.) - 1.);
        s0 += (x0 < 700.) ? b0 : 0.;
    }
    return s0 + s1;
}

// The adaptive walk of planck_nu1_nu2 over the scaled integrand nu^3 / (exp(a nu) - 1), a = (h/k_B) / T (B_nu without its constant
// factor 2h/c^2, which planck_nu1_nu2 applies once to the finished integral), written as nu^3 w / (1 - w) with w = exp(-a nu):
//   * the walk is FP64-pipe-bound (profiles/r02_lx_pass_experiment.md: 143 FP64 instructions per trial step, 96 of them in the
//     four exponentials of the stages), and the four stage abscissae of a Cash-Karp step sit at t + {12, 24, 40, 35}/40 dt: ONE
//     exponential v = exp(-a dt / 40) and a product chain give all four factors (v^12, v^24, v^35, v^40), and w at the new t
//     is w_t v^40 — 24 FP64 instructions instead of 56.  A product of <= 40 factors carries <= 50 ulp = 5.5e-15 relative, w_t
//     collects that once per accepted step (a walk has ~10): 1e-13 at most, three orders below the 1e-10 parity tolerance and
//     far below what could flip a controller decision (|err - threshold| would have to be < 1e-13 of the threshold);
//   * where w < 2^-54 (a nu > 37.4: most of the band for most rings) 1 - w is exactly 1 and the reciprocal drops out — decided
//     per warp, and the bits do not depend on the decision (fb_rcp(1) is exactly 1);
//   * exp(-a nu) underflows smoothly, so the walk needs no overflow guards (the argument of the one direct exponential is held at
//     -700, fb_exp_tab's range: e^-700 against the hottest ring's e^-1 is zero at any precision).
FB_HD double planck_walk(double hkT, double nu1, double nu2, double tol, int* trials) {
    const double b1 = FB_KC(CK_B1), b3 = FB_KC(CK_B3), b4 = FB_KC(CK_B4), b6 = FB_KC(CK_B6);
    const double db1 = FB_KC(CK_DB1), db3 = FB_KC(CK_DB3), db4 = FB_KC(CK_DB4), db5 = FB_KC(CK_DB5), db6 = FB_KC(CK_DB6);
    const double c3 = FB_KC(CK_C3), c4 = FB_KC(CK_C4), c6 = 7. / 8.;
    const double eps_mach = 2.220446049250313e-16;
    const double tiny_w = 5.551115123125783e-17;  // 2^-54: 1 - w == 1 below
    double x = 0., t = nu1, dt = 1e-2 * (nu2 - nu1);
    int fails = 0, n = 0;
    double wt;  // exp(-a t)
    {
        const double a = hkT * t;
        wt = fb_exp_tab(-((a < 700.) ? a : 700.));
    }
    double k1 = (p3(t) * wt) * fb_rcp(1. - wt);
    while (nu2 - t > eps_mach) {
        if ((t + dt) - nu2 > eps_mach) dt = nu2 - t;  // less_with_sign(end, t + dt, dt)
        // stage 2 has zero weight in both the solution and the error estimate (b2 = db2 = 0) and the integrand does not depend
        // on the state: it is not evaluated
        double w3, w4, w5, w6;
        {
            const double u = (hkT * dt) * 0.025;
            const double v = fb_exp_tab(-((u < 700.) ? u : 700.));
            const double v2 = v * v, v3 = v2 * v, v5 = v3 * v2, v6 = v3 * v3, v11 = v6 * v5, v12 = v6 * v6, v24 = v12 * v12;
            const double v35 = v24 * v11;
            w3 = wt * v12; w4 = wt * v24; w6 = wt * v35; w5 = wt * (v35 * v5);
        }
        bool no_rcp = wt < tiny_w;  // every stage's w is below w_t
#ifdef __CUDA_ARCH__
        no_rcp = __all_sync(__activemask(), no_rcp);
#endif
        double k3 = p3(t + c3 * dt) * w3, k4 = p3(t + c4 * dt) * w4, k5 = p3(t + dt) * w5, k6 = p3(t + c6 * dt) * w6;
        if (!no_rcp) {
            k3 *= fb_rcp(1. - w3); k4 *= fb_rcp(1. - w4); k5 *= fb_rcp(1. - w5); k6 *= fb_rcp(1. - w6);
        }
        ++n;
        // odeint forms x + (b1 dt) k1 + (b3 dt) k3 + ...; with dt factored out the two sums are 11 instead of 19 FP64
        // instructions and differ from that order by rounding only.
        const double xnew = fma(dt, fma(b6, k6, fma(b4, k4, fma(b3, k3, b1 * k1))), x);
        const double xerr = dt * fma(db6, k6, fma(db5, k5, fma(db4, k4, fma(db3, k3, db1 * k1))));
        const double err = fabs(xerr) * fb_rcp14(tol * (fabs(x) + fabs(dt) * fabs(k1)));  // feeds comparisons and the step factor only
        // Step-size controller.  Both adjustments are powers of the error, so one power evaluation serves the
        // rejecting and the accepting lanes of a warp alike:
        //   reject (err > 1):      dt *= max(0.9 err^(-1/3), 0.2)
        //   accept, err < 0.5:     dt *= 0.9 max(5^-5, err)^(-1/5)
        const bool reject = err > 1.0;
        const double floor_ = FB_KL(CK_FLOOR);  // pow(5.0, -5)
        double fac;
        if (!reject && !(floor_ < err)) {
            fac = 4.5;  // 0.9 * (5^-5)^(-1/5): the growth limit — the common case on the exponential tail of the band
        } else {
            // (a rejected step with err >= 91.125 = (0.9 / 0.2)^3 gets the floor 0.2 below whatever the power: err is held at 100)
            fac = FB_KL(CK_NINE_TENTHS) * fb_pow_step_factor((err < 100.) ? err : 100., reject);
        }
        if (reject) {
            dt *= (fac < FB_KL(CK_FIFTH)) ? FB_KL(CK_FIFTH) : fac;
            if (++fails >= 500) break;  // odeint throws here; unreachable for a Planck integrand
            continue;
        }
        fails = 0;
        t += dt;  // == the abscissa of stage 5 (c5 = 1): the derivative at the new t is k5, and w there is w5
        if (err < 0.5) dt *= fac;
        x = xnew;
        k1 = k5;
        wt = w5;
    }
    if (trials) *trials = n;
    return x;
}

// integral of B_nu over [nu1, nu2]: boost::numeric::odeint::integrate_adaptive with
// make_controlled(eps_abs = 0, eps_rel = tol, runge_kutta_cash_karp54<double>), initial step
// 0.01 (nu2 - nu1) (cpp/src/spectrum.cpp:26-37).  Controller = odeint's default error checker
// (a_x = a_dxdt = 1) and default step adjuster: reject when err > 1 with dt *= max(0.9 err^(-1/3), 0.2);
// on acceptance grow when err < 0.5 with dt *= 0.9 max(5^-5, err)^(-1/5).  `trials` counts try_step calls.
// The constant factor 2h/c^2 of B_nu is applied to the finished integral (the controller's error ratio does not see a
// common factor of the integrand), and the walk is compiled twice: with and without the exp-overflow range checks.
FB_HD double planck_nu1_nu2(double T, double nu1, double nu2, double tol, int* trials) {
    if (trials) *trials = 0;
    // Identically-zero integrand: T <= 0 or NaN (Planck_nu returns 0), or exp() overflows to +inf already at
    // nu1 so every sample is h nu^3 / inf = 0.  The reference then accepts 100 constant steps of zeros
    // (error 0/0 = NaN compares false both ways) and returns exactly 0; skip the walk.
    if (!(T > 0.)) return 0.;
    const double hkT = FB_H_OVER_KB * fb_rcp(T);
    if (hkT * nu1 > 710.) return 0.;
    return FB_DOUBLE_H_OVER_C2 * planck_walk(hkT, nu1, nu2, tol, trials);
}

// The same integral through the table of the walk's result (lx_table.h): nu1^4 q(x1) e^{-x1}, x1 = h nu1 / kT — one look-up,
// one exponential and a degree-11 polynomial, <= 1e-12 from the walk, with the walk's own try_step count.  `tab` = nullptr, or
// x1 outside the table: the walk.
FB_HD double planck_nu1_nu2_tab(const double* __restrict__ tab, double T, double nu1, double nu2, double tol, int* trials) {
    *trials = 0;
    if (!(T > 0.)) return 0.;
    const double hkT = FB_H_OVER_KB * fb_rcp(T);
    const double x1 = hkT * nu1;
    if (x1 > 710.) return 0.;
    if (tab != nullptr && x1 >= FB_LX_X_LO && x1 < FB_LX_X_HI) {
        const double q = lx_table_q(tab, x1, trials);
        if (*trials >= 0) return (FB_DOUBLE_H_OVER_C2 * p4(nu1)) * (q * fb_exp_tab(-x1));
        // a sliver around a breakpoint of the walk (a few 1e-9 wide; a ring lands in one about once in 1e8): walk
    }
    return FB_DOUBLE_H_OVER_C2 * planck_walk(hkT, nu1, nu2, tol, trials);
}

// ------------------------------------------------------------------------------------------------
// per-point set-up
// ------------------------------------------------------------------------------------------------
// Geometric factor of T_GR (cpp/src/spectrum.cpp:44-68) that depends on the radius only:
//   T_GR^4 sigma = (3 Mdot c^6 / (8 pi GM^2)) * Gb(r).   Returns -1 for x < x0 (T_GR = 0 there).
FB_HD double tgr_geom(const fb200_model_dev_t& M, double r1, int* inside) {
    const double ak = M.kerr;
    const double x = sqrt(r1 / M.tgr_rg);
    const double x0 = M.tgr_x0;
    if (x < x0) {
        *inside = 1;
        return 0.;
    }
    *inside = 0;
    const double x1 = M.tgr_x1, x2 = M.tgr_x2, x3 = M.tgr_x3;
    const double a = 3. * (x1 - ak) * (x1 - ak) * log((x - x1) / (x0 - x1)) / x1 / (x1 - x2) / (x1 - x3);
    const double b = 3. * (x2 - ak) * (x2 - ak) * log((x - x2) / (x0 - x2)) / x2 / (x2 - x1) / (x2 - x3);
    const double c = 3. * (x3 - ak) * (x3 - ak) * log((x - x3) / (x0 - x3)) / x3 / (x3 - x1) / (x3 - x2);
    return (x - x0 - 1.5 * ak * log(x / x0) - a - b - c) / (p4(x) * (p3(x) - 3. * x + 2. * ak));
}

// ------------------------------------------------------------------------------------------------
// dynamic winds: value of the wind term at point i for the current Mdot_in (cpp/src/freddi_state.cpp:457-654)
// Returns C (and B through *Bout for Janiuk2015).  `inside` = first <= i <= last.
// ------------------------------------------------------------------------------------------------
struct WindScalars {
    double v0, v1, v2, v3, v4;
};

FB_HD WindScalars wind_scalars(const fb200_model_dev_t& M, double Mdot, double h_first, double h_last, double h_back) {
    WindScalars w{0, 0, 0, 0, 0};
    const double mu = M.mu;
    const double c2 = p2(FB_C_LIGHT);
    const double L = Mdot * c2 * M.eta;
    const double L_edd = (4.0 * FB_PI * M.GM * 2.0 * mu * FB_M_PROTON * FB_C_LIGHT / FB_THOMSON);
    switch (M.windtype) {
        case FB200_WIND_SHIELDSOSCIL1986:
            w.v0 = sqrt(M.windparams[1]) * h_back;               // h_wind_min
            w.v1 = -0.5 / FB_PI * M.windparams[0] * Mdot;         // prefactor
            w.v2 = log(1 / M.windparams[1]);
            break;
        case FB200_WIND_JANIUK2015: {
            w.v0 = 2 * M.GM / c2;  // R_g
            const double lol = L / L_edd;
            w.v1 = 1 - 1 / (1 + M.windparams[0] * p2(lol));  // fout
            break;
        }
        case FB200_WIND_SHIELDS1986:
        case FB200_WIND_WOODS1996: {
            const double T_ic = M.windparams[1];
            w.v0 = (M.GM * mu * FB_M_PROTON) / (FB_K_BOLTZ * T_ic);  // R_iC
            const double L_crit = (1.0 / 8.0) * sqrt(FB_M_ELECTRON / (mu * FB_M_PROTON)) *
                                  sqrt((FB_M_ELECTRON * FB_C_LIGHT * FB_C_LIGHT) / (FB_K_BOLTZ * T_ic)) * L_edd;
            w.v1 = L / L_crit;  // el
            w.v2 = L;
            w.v3 = pow(w.v1, 2.0 / 3.0);
            break;
        }
        case FB200_WIND_WOODS1996AGN: {
            const double T_ic = M.windparams[1];
            const double R_iC = (M.GM * mu * FB_M_PROTON) / (FB_K_BOLTZ * T_ic);
            const double le = L / L_edd;
            double R_tr;
            if (le <= 0.01) R_tr = 6.0 * R_iC;
            else R_tr = R_iC * (6.0 + 5.4 * log(le / 0.01) + 4.1 * ((log(le / 0.01)) * (log(le / 0.01))));
            double f_L;
            if (le <= 0.1) f_L = 1.0;
            else f_L = pow((0.1 / le), 0.15);
            w.v0 = R_iC; w.v1 = le; w.v2 = R_tr; w.v3 = f_L;
            break;
        }
        case FB200_WIND_TOY:
            w.v0 = Mdot; w.v1 = h_first; w.v2 = p2(h_last - h_first);
            break;
        default: break;
    }
    return w;
}

FB_HD void wind_dynamic_point(const fb200_model_dev_t& M, const WindScalars& w, double h, double R, double* Bout, double* Cout) {
    const double GM = M.GM;
    switch (M.windtype) {
        case FB200_WIND_SHIELDSOSCIL1986:
            if (h > w.v0) *Cout = w.v1 / (w.v2 * p2(R)) * (4 * FB_PI * p3(h)) / p2(GM);
            break;
        case FB200_WIND_JANIUK2015:
            if (R > 70.0 * w.v0) {
                const double C_0 = (4.0 * FB_PI * p3(h)) / (p2(GM));
                const double Q = 2 * (3 / (8 * FB_PI)) * ((p4(GM)) / (p7(h)));
                *Bout = -0.75 * C_0 * (1 / M.windparams[1]) * ((4 * Q * R) / (3 * GM)) * w.v1;
            }
            break;
        case FB200_WIND_SHIELDS1986: {
            const double R_iC = w.v0, el = w.v1, L = w.v2;
            if (R > 0.1 * R_iC) {
                const double Xi_max = M.windparams[0], T_ic = M.windparams[1], Pow = M.windparams[2];
                const double xi = R / R_iC;
                const double T_ch = T_ic * w.v3 * pow(xi, -2.0 / 3.0);
                const double P_0 = L / (4.0 * FB_PI * p2(R) * Xi_max * FB_C_LIGHT);
                const double C_ch = sqrt((FB_K_BOLTZ * T_ch) / (M.mu * FB_M_PROTON));
                const double m_ch0 = P_0 / C_ch;
                const double C0 = (4.0 * FB_PI * p3(h)) / (p2(GM));
                const double q = 1.2 * xi / (xi + el) + 2.2 / (1. + p2(el) * xi);
                const double y = sqrt(1. + 1. / (4. * p2(xi)) + ((p2(xi)) / (1. + p2(xi))) * (q * q));
                const double Mach_cc = cbrt(((1. + (el + 1.) / xi) / (1. + 1. / ((1. + p2(xi)) * p4(el)))));
                const double p_po = 0.5 * exp(-(((1. - 1. / y) * (1. - 1. / y)) / (2. * xi)));
                *Cout = -2.0 * Pow * C0 * m_ch0 * Mach_cc * p_po * p2(y);
            }
            break;
        }
        case FB200_WIND_WOODS1996AGN: {
            const double R_iC = w.v0, le = w.v1, R_tr = w.v2, f_L = w.v3;
            const double g_R = (R <= R_tr) ? 1.0 : R / R_tr;
            const double xi1 = R_iC / R;
            const double C0 = (4.0 * FB_PI * p3(h)) / (p2(GM));
            const double s = 1.0 - (1 / sqrt(1.0 + 0.25 * p2(xi1)));
            const double ExP = exp(-((s * s) / (2.0 / xi1)));
            *Cout = -2.0 * (2e42 / (M.Mx)) * (M.windparams[0] / 1e13) * C0 * le * p2(xi1) * f_L * g_R * ExP;
            break;
        }
        case FB200_WIND_WOODS1996: {
            const double R_iC = w.v0, el = w.v1, L = w.v2;
            if (R > 0.1 * R_iC) {
                const double Xi_max = M.windparams[0], T_ic = M.windparams[1], Pow = M.windparams[2];
                const double xi = R / R_iC;
                const double xi1 = R_iC / R;
                const double T_ch = T_ic * w.v3 * pow(xi, -2.0 / 3.0);
                const double C_ch = sqrt((FB_K_BOLTZ * T_ch) / (M.mu * FB_M_PROTON));
                const double C0 = (4.0 * FB_PI * p3(h)) / (p2(GM));
                const double Fr = L / (4.0 * FB_PI * p2(R) * Xi_max * C_ch * FB_C_LIGHT);
                const double Fc = ((pow((1 + p2(((0.125 * el + 0.00382) * xi1))), (1.0 / 6.0))) /
                                   pow((1 + 1. / p2((p4(el) * (1.0 + 262.0 * p2(xi))))), (1.0 / 6.0)));
                const double s = 1.0 - (1 / sqrt(1.0 + 0.25 * p2(xi1)));
                const double Expo = exp(-((s * s) / (2.0 * xi)));
                *Cout = -2.0 * Pow * C0 * Fr * Fc * Expo;
            }
            break;
        }
        case FB200_WIND_TOY: {
            const double Mdot = w.v0 * (h - w.v1) / w.v2;
            *Cout = -2.0 * M.windparams[0] * Mdot;
            break;
        }
        default: break;
    }
}

// ------------------------------------------------------------------------------------------------
// scalars of the current state that several getters share
// ------------------------------------------------------------------------------------------------
struct StateScalars {
    double Mdot_plain;  // (F[first+1]-F[first])/(h[first+1]-h[first])      FreddiState::Mdot_in, freddi_state.cpp:156-158
    double Mdot;        // virtual Mdot_in(): + dFmagn_dh[first] for NS       ns_evolution.cpp:471-474
    double Lbol_disk;   // freddi_state.cpp:161-163 / ns_evolution.cpp:526-528
    double Lbol_ns_rest, Lbol_ns, eta_ns, fp, R_alfven;  // ns_evolution.cpp:386-395,476-478
};

FB_HD double ns_R_alfven(const fb200_model_dev_t& M, double Mdot) {  // ns_arguments.cpp:50-52
    return M.epsilon_Alfven * pow(p4(M.mu_magn) / (M.GM * p2(Mdot)), 1. / 7.);
}

template <class Ex>
FB_HD StateScalars state_scalars(Ex& ex, const fb200_model_dev_t& M, const typename Ex::Arrays& S, int first, bool want_ns_lum) {
    StateScalars s{};
    const double F0 = S.F[first], F1 = S.F[first + 1], h0 = S.h[first], h1 = S.h[first + 1];
    s.Mdot_plain = (F1 - F0) / (h1 - h0);
    s.Mdot = s.Mdot_plain;
    if (M.is_ns) {
        s.Mdot = s.Mdot_plain + ex.dFmagn_dh(first);
        const double R0 = S.R[first];
        s.Lbol_disk = (F0 + 0.5 * s.Mdot * h0) * sqrt(M.GM / (R0 * R0 * R0));
        if (want_ns_lum) {
            s.R_alfven = ns_R_alfven(M, s.Mdot);
            s.eta_ns = ns_eta(M, s.R_alfven, R0);
            s.fp = ns_fp(M, R0);
            double Mdot = s.Mdot;
            if (Mdot < 0.0) Mdot = 0.0;
            s.Lbol_ns_rest = s.eta_ns * s.fp * Mdot * p2(FB_C_LIGHT);
            s.Lbol_ns = M.redshift * s.Lbol_ns_rest;
        }
    } else {
        s.Lbol_disk = M.eta * s.Mdot * p2(FB_C_LIGHT);
    }
    return s;
}

// ------------------------------------------------------------------------------------------------
// hot-zone structure at one point (cpp/src/freddi_state.cpp:184-290, opacity_related.cpp:85-87)
// ------------------------------------------------------------------------------------------------
struct PointStructure {
    double W, Sigma, Height, Kirr, Qx, Tvis, Tph, Tirr;
};

FB_HD double pow025(double x) { return sqrt(sqrt(x)); }

// W = |F|^(1-m) h^n / (1-m) / D (freddi_evolution.cpp:80); hnD = h^n / (1-m) / D is a per-ring constant
// (init_points), so one multiplication replaces the two divisions (re-association: <= 1 ulp).
FB_HD double wunc_point(const fb200_model_dev_t& M, double F, double hnD) {
    return fb_pow_pos(fabs(F), M.one_minus_m) * hnD;
}

// W_known >= 0: the closure value is already known (see structure_and_row); need_height = false: Height (a pow per ring) is
// neither output nor used — isotropic emitters with irrindex = 0 make Qx independent of H/R — and is left 0.
FB_HD PointStructure hot_structure(const fb200_model_dev_t& M, const StateScalars& sc, double h, double R, double hn, double hm175,
                                   double hotR, double F, double W_known = -1., bool need_height = true) {
    PointStructure q;
    q.W = (W_known >= 0.) ? W_known : wunc_point(M, F, hn);
    q.Sigma = q.W * p2(M.GM) / (4. * FB_PI * p3(h));
    q.Height = need_height ? hotR * fb_pow_pos(F, M.Height_exp_F) : 0.;
    const double mu = q.Height / R;
    q.Kirr = M.Cirr * ((M.irrindex == 0.) ? 1.0 : pow(q.Height / (R * 0.05), M.irrindex));
    if (M.is_ns) {
        q.Qx = q.Kirr * (sc.Lbol_disk * angular_dist(M.angdist_disk, mu) + sc.Lbol_ns * angular_dist(M.angdist_ns, mu)) /
               (4. * FB_PI * p2(R));
    } else {
        q.Qx = q.Kirr * sc.Lbol_disk * angular_dist(M.angdist_disk, mu) / (4. * FB_PI * p2(R));
    }
    q.Tvis = hm175 * pow025(3. / (8. * FB_PI) * F / FB_SIGMA_SB);
    q.Tph = pow025(p4(q.Tvis) + q.Qx / FB_SIGMA_SB);
    q.Tirr = pow025(q.Qx / FB_SIGMA_SB);
    return q;
}

// The same quantities for the row's structure pass, where only Sigma, Tph and Tirr of a ring are needed (Tvis, Kirr and
// H/R are output at the outer edge only, which uses hot_structure): divisions by per-ring denominators become
// reciprocal-multiplies, Tvis^4 is formed without taking the fourth root first, Tirr of an unirradiated ring is 0 without
// a square root.  ~110 instead of ~250 FP64 instructions per ring; differs from hot_structure by rounding only.
struct RowStructure {
    double Sigma, Tph, Tirr, H2R;
};
// sgA = GM^2 / (4 pi h^3), qB = 1 / (4 pi R^2 sigma_SB), tv4 = (GM h^-1.75)^4: the grid-only factors, set up once per model
// (init_points) instead of two reciprocals and a fourth power per ring and row.
FB_HD RowStructure hot_structure_lean(const fb200_model_dev_t& M, const StateScalars& sc, double R, double hn, double sgA, double qB, double tv4,
                                      double hotR, double F, double W_known, bool need_height) {
    RowStructure q;
    const double W = (W_known >= 0.) ? W_known : wunc_point(M, F, hn);
    q.Sigma = W * sgA;
    const double mu = need_height ? hotR * fb_pow_pos(F, M.Height_exp_F) * fb_rcp(R) : 0.;
    q.H2R = mu;
    const double Kirr = M.Cirr * ((M.irrindex == 0.) ? 1.0 : pow(mu / 0.05, M.irrindex));
    double L = sc.Lbol_disk * angular_dist(M.angdist_disk, mu);
    if (M.is_ns) L += sc.Lbol_ns * angular_dist(M.angdist_ns, mu);
    const double Qs = Kirr * L * qB;  // Qx / sigma_SB
    const double Tvis4 = tv4 * ((3. / (8. * FB_PI) / FB_SIGMA_SB) * F);
    q.Tph = pow025(Tvis4 + Qs);
    q.Tirr = (Qs == 0.) ? 0. : pow025(Qs);
    return q;
}

// Sigma_minus / Sigma_plus / v_cooling_front (cpp/src/freddi_state.cpp:692-722)
FB_HD double sigma_minus(const fb200_model_dev_t& M, double r) {
    return 74.6 * M.sigma_minus_k * pow(r / 1e10, 1.18) * M.sigma_minus_mx;
}
FB_HD double sigma_plus(const fb200_model_dev_t& M, double r) {
    return 39.9 * M.sigma_plus_k * pow(r / 1e10, 1.11) * M.sigma_plus_mx;
}
FB_HD double v_cooling_front(const fb200_model_dev_t& M, double r, double Sigma_last) {
    const double Sigma_plus_ = sigma_plus(M, r);
    const double sigma = log(Sigma_last / Sigma_plus_) / log(sigma_minus(M, r) / Sigma_plus_);
    return 1e5 * (1.439 - 5.305 * sigma + 10.440 * p2(sigma) - 10.55 * p3(sigma) + 4.142 * p4(sigma)) *
           pow(M.alpha / 0.2, 0.85 - 0.69 * sigma) * pow(M.alphacold / 0.05, 0.05 + 0.69 * sigma) * pow(r / 1e10, 0.035) *
           M.vcf_mx;
}


// Weight of point i in trapz(x, y, first, last) without its factor 1/2 (cpp/src/util.cpp:9-19): one-sided at the two ends.
// The kernels take every radial integral (Mdisk, Lx, Fnu, Mdot_wind) as a CTA-wide sum of y_i times this weight.
template <class X>
FB_HD double trapz_weight(const X& x, int i, int first, int last) {
    if (i == first) return x[i + 1] - x[i];
    if (i == last) return x[i] - x[i - 1];
    return x[i + 1] - x[i - 1];
}

// trapz(x, y, first, last) of cpp/src/util.cpp:9-19 the way the kernels evaluate it: reduce_sums over the points.
template <class Ex, class X, class Y>
FB_HD double policy_trapz(Ex& ex, const X& x, const Y& y, int first, int last) {
    if (first >= last) return 0.;
    ex.reduce_sums(1, [&](int i, int k, int q) -> double { return (i < first || i > last) ? 0. : y[i] * trapz_weight(x, i, first, last); });
    return 0.5 * ex.sum(0);
}

// The tridiagonal solve is block-partitioned: FB200_BLK consecutive unknowns are eliminated sequentially by
// one thread (Thomas-style, in registers), which leaves a tridiagonal "reduced" system in the first and last
// unknown of every block (2 per block, <= FB200_NRED_MAX) that is solved by parallel cyclic reduction.
#define FB200_BLK 8
#define FB200_NRED_MAX (2 * (FB200_NX_MAX / FB200_BLK))
// XOR swizzle inside aligned groups of 8: a thread walking its own block (stride-8 words across the lanes of a
// warp) then hits 16 distinct 8-byte bank pairs per half-warp instead of 2, while a warp reading 32 consecutive
// points still covers 32 consecutive words.  No padding needed.
FB_HD int pidx(int i) { return i ^ ((i >> 4) & 7); }
#define FB200_NPAD FB200_NX_MAX

FB_HD bool fb_isfinite(double x) { return fabs(x) <= 1.7976931348623157e308; }
FB_HD bool fb_isnan(double x) { return x != x; }
// The NaN an invalid operation produces on the reference's platform (x86-64 SSE2: the default QNaN has the sign bit
// set; glibc's pow(negative, 0.25) returns (x - x) / (x - x)).  The value is the same NaN either way; the sign only
// shows in text output ("-nan"), which the freddi.dat writer reproduces byte for byte.
#define FB_NAN_INVALID (-NAN)

// ------------------------------------------------------------------------------------------------
// The engine.  Execution policy `Ex` provides:
//   points(f)            run f(i) for the point(s) this thread owns (no barrier)
//   sync()               CTA barrier
//   reduce_max(f)        max over points of f(i), NaN-skipping (fmax), value returned to every thread
//   reduce_max_int(f) / reduce_min_int(f)
//   reduce_sums(nq, f)   sum over points of f(i, q) for q < nq; results via sum(q)
//   scalar(k)            small shared scratch (written by one thread, read after sync)
//   windA(i), windB(i), windC_static(i), Fmagn(i), dFmagn_dh(i)   per-model arrays in global memory (0 if absent)
//   lambda_pref(k)       double_h_c2 / lambda_k^5
//   blocks(n, f)         run f(b) for b < n, one thread per b (the in-thread block elimination of the solve)
//   reduced(n, f)        run f(q) for q < n, one thread per q (reduced-system PCR)
//   count(c, v)          add v (a uniform value) to bookkeeping counter c (FB_CNT_*)
//   note_pow_fallbacks(n)  n rings of this thread left the binomial series of the incremental K update (bookkeeping only)
//   note_wind_update(Mdot, first, last)   arguments of the dynamic wind's update() of this step (read-out of windB/windC/Mdot_wind)
//   tick(p)              phase boundary marker (cycle accounting in FB_PROFILE builds, otherwise a no-op)
//   tri_points(n, f)     run f(t) for t < n (the triangles of the companion star), spread over the threads
//   reduce_sums3(n, f)   sums over t < n of the three values f(t, a, b, c) produces; results via sum(0..2)
//   star_T(t), star_sum(k)   per-triangle effective temperature; the 3 x (n_lambdas + n_passbands) star sums
// ------------------------------------------------------------------------------------------------
// kExtras = false compiles the passband and companion-star columns out (they cost registers and ~2 % of the kernel
// even when unused): the evolve kernel is instantiated both ways and launched by what the ensemble asks for.
template <class Ex, bool kExtras = true>
struct DiscEngine {
    Ex& ex;
    const fb200_model_dev_t& M;
    const OutCfg& cfg;
    typename Ex::Arrays S;
    ModelState st;  // uniform copy (every thread holds the same values)
    bool k_valid = false;  // S.s_K holds the closure of the state in S.F for every interior ring (set by a completed solve)

    FB_HD DiscEngine(Ex& ex_, const fb200_model_dev_t& M_, const OutCfg& cfg_, const typename Ex::Arrays& S_, const ModelState& st_)
        : ex(ex_), M(M_), cfg(cfg_), S(S_), st(st_) {}

    // storage position of reduced unknown q (see solve_step)
    FB_HD static int rpos(int q) { return Ex::kWarpPcr ? ((q & 7) * 32 + (q >> 3)) : q; }
#ifdef FB_PCR_ODD
    // Odd-only reduced solve (solve_step, -DFB_PCR_ODD: measured equal in time to the full PCR): where block b's two
    // reduced rows are stored in buffer 0.  Plain policies: q = 2b (the
    // row of the block's first unknown, "even") and 2b + 1 ("odd").  Hybrid (warp) solver: the odd rows in the lower half of the
    // buffer in the warp-major order of its four warps (unknown p = warp + 4 lane at position 32 warp + lane), the even rows
    // likewise in the upper half.
    FB_HD static int ppos(int b) { return (b & 3) * 32 + (b >> 2); }
    FB_HD static int epos(int b) { return Ex::kWarpPcr ? FB200_NRED_MAX / 2 + ppos(b) : 2 * b; }
    FB_HD static int opos(int b) { return Ex::kWarpPcr ? ppos(b) : 2 * b + 1; }
#else
    FB_HD static int epos(int b) { return rpos(2 * b); }
    FB_HD static int opos(int b) { return rpos(2 * b + 1); }
#endif

    // nonlinear_diffusion.cpp:46-51 for an interior point i (1 <= i <= Nx-2), wind terms A, B
    FB_HD void interior_coeffs(int i, double A, double B) const {
        const double xm = S.h[i - 1], x = S.h[i], xp = S.h[i + 1];
        S.ca[i] = (xp - x) / (xp - xm) * (2.0 - A * (xp - x));
        S.cb[i] = (x - xm) / (xp - xm) * (2.0 + A * (x - xm));
        S.cc[i] = 2.0 - A * (xp - 2 * x + xm) - B * (xp - x) * (x - xm);
        S.frac[i] = (xp - x) * (x - xm) / M.tau;
    }

    // ---- one-time per model: grid-derived constants.  S.h and S.F must be loaded and synced. ----
    FB_HD void init_points() {
        const int Nx = M.Nx;
        ex.points([&](int i, int k) {
            if (i >= Nx) return;
            const double h = S.h[i];
            const double R = p2(h) / M.GM;  // FreddiState::DiskStructure::initialize_R, freddi_state.cpp:47-53
            S.R[i] = R;
            S.hn[i] = pow(h, M.n_op) / M.one_minus_m / M.D;
            S.c1[i] = 3. / (8. * FB_PI) * p4(M.GM) / p7(h);                                      // freddi_state.cpp:300
            S.hm175[i] = M.GM * pow(h, -1.75);                                                    // freddi_state.cpp:284
            S.sgA[i] = p2(M.GM) * fb_rcp(4. * FB_PI * p3(h));                                     // Sigma = W GM^2 / (4 pi h^3), :198
            S.qB[i] = fb_rcp(4. * FB_PI * p2(R)) * (1. / FB_SIGMA_SB);                            // Qx / (Kirr L sigma_SB), :240
            S.tv4[i] = p4(S.hm175[i]);
            S.hotR[i] = R * M.Height_coef * pow(R / 1e10, M.Height_exp_R - M.Height_exp_F / 2.);  // opacity_related.cpp:86
            int inside;
            const double g = tgr_geom(M, R, &inside);
            S.Gb[i] = inside ? 0. : g;
            if (i > 0 && i < Nx - 1) interior_coeffs(i, ex.windA(i), ex.windB(i));
            else { S.ca[i] = 0.; S.cb[i] = 0.; S.cc[i] = 0.; S.frac[i] = 0.; }
        });
        ex.sync();
    }

    // ---- truncateInnerRadius, NS only (ns_evolution.cpp:431-468).  Returns false on RadiusCollapse. ----
    FB_HD bool truncate_inner() {
        if (M.R_dead <= 0.) return true;
        const StateScalars sc = state_scalars(ex, M, S, st.first, false);
        const double Mdot = sc.Mdot;
        if ((!fb_isfinite(st.Mdot_in_prev)) || (Mdot < 0.0) || (st.Mdot_in_prev < 0.0)) return true;
        if (Mdot > st.Mdot_in_prev) return true;
        const double Ra = ns_R_alfven(M, Mdot);
        double R_m = (M.R_m_min < Ra) ? Ra : M.R_m_min;  // std::max(R_m_min, R_Alfven)
        R_m = (M.R_dead < R_m) ? M.R_dead : R_m;         // std::min(R_m, R_dead)
        const int first = st.first, last = st.last;
        // smallest ii in [first, last-2] with R[ii+1] > R_m, else the loop runs out at last-1
        int ii = ex.reduce_min_int([&](int i, int k) -> int {
            if (i >= first && i <= last - 2 && S.R[i + 1] > R_m) return i;
            return 0x7fffffff;
        });
        if (ii == 0x7fffffff) ii = last - 1;
        if (ii >= last - 2) return false;
        R_m = S.R[ii];
        st.first = ii;
        double new_F_in = 0;
        if (M.inverse_beta <= 0.) {
            if (R_m <= M.R_cor) new_F_in = kappa_t(M, R_m) * p2(M.mu_magn) / p3(M.R_cor);
            else new_F_in = kappa_t(M, R_m) * p2(M.mu_magn) / p3(R_m);
        }
        st.F_in = new_F_in;
        return true;
    }

    // ---- the implicit step (nonlinear_diffusion.cpp:31-89) ----
    // On entry S.F holds y(t); on exit y(t+tau).  Returns Picard iterations, or -1 if the guard tripped.
    FB_HD int solve_step(double Mdot_in_for_wind) {
        const int first = st.first, last = st.last;
        const int m = last - first;  // unknowns first+1..last
        const int Nx = M.Nx;
        const double tau = M.tau, F_in = st.F_in, rbc = M.Mdot_out;
        WindScalars ws{0, 0, 0, 0, 0};
        if (M.wind_dynamic) {
            ws = wind_scalars(M, Mdot_in_for_wind, S.h[first], S.h[last], S.h[Nx - 1]);
            ex.note_wind_update(Mdot_in_for_wind, first, last);  // what this update() saw
        }
        // closure pow() calls of this step (roofline bookkeeping): the outer ring always; the interior rings unless K is carried
        ex.count(FB_CNT_POW_EVALS, k_valid ? 1 : m);
        const bool janiuk = (M.windtype == FB200_WIND_JANIUK2015);
        // the initial K holds the source term tau C (nonlinear_diffusion.cpp:61): only without one is it the pure closure
        const bool has_C = M.wind_dynamic || ex.has_windC();

        if (janiuk) {  // B changes with Mdot_in: refresh the interior coefficients
            ex.points([&](int i, int k) {
                if (i > 0 && i < Nx - 1) {
                    double Bd = 0., Cd = 0.;
                    if (i >= first && i <= last) wind_dynamic_point(M, ws, S.h[i], S.R[i], &Bd, &Cd);
                    interior_coeffs(i, 0., Bd);
                }
            });
            // each thread rewrites only its own entries and reads them back itself: no barrier needed
        }

        ex.points([&](int i, int k) {
            if (!(i > first && i <= last)) return;
            const bool is_last = (i == last);
            double K;
            const double h = S.h[i], R = S.R[i];
            double A = ex.windA(i), B = ex.windB(i), C = ex.windC_static(i);
            if (M.wind_dynamic) {
                double Bd = 0., Cd = 0.;
                wind_dynamic_point(M, ws, h, R, &Bd, &Cd);  // update() rewrites the points of [first, last]
                if (janiuk) B = Bd; else C = Cd + C;
            }
            const double y = S.F[i];
            double lo, up, cc, frac, rhs;
            if (!is_last && k_valid) {
                // An interior ring of this step was one in the previous step too (`first` only grows, `last` only shrinks), and
                // s_K still holds K = frac W(y) / y of that step's last iteration for the very y in S.F: frac W = K y, no pow.
                lo = S.ca[i]; up = S.cb[i]; cc = S.cc[i]; frac = S.frac[i];
                const double Kp = S.s_K[i];
                const double fW = Kp * y;
                if (C == 0.) { K = Kp; rhs = fW; }
                else { rhs = fW + frac * (tau * C); K = rhs / y; }
            } else if (!is_last) {
                const double W = wunc_point(M, y, S.hn[i]);
                lo = S.ca[i]; up = S.cb[i]; cc = S.cc[i]; frac = S.frac[i];
                const double f = frac * (W + tau * C);
                K = f / y;                // nonlinear_diffusion.cpp:61 — includes tau*C
                rhs = f;
            } else {
                const double W = wunc_point(M, y, S.hn[i]);
                const double d = h - S.h[i - 1];
                const double aL = 1 - 0.5 * A * d;
                lo = aL; up = 0.;
                cc = aL - 0.5 * B * d * d;
                frac = d * d * 0.5 / tau;
                const double f = frac * (W + tau * C);
                K = frac * W / y;         // nonlinear_diffusion.cpp:65 — no C, never updated in the loop
                rhs = d * rbc + f;
            }
            if (i == first + 1) { rhs += lo * F_in; lo = 0.; }  // y[first] = F_in is known
            const int pi = pidx(i);
            S.s_l[pi] = -lo; S.s_u[pi] = -up; S.s_rhs[pi] = rhs; S.s_d[pi] = cc + K;
            S.s_cc[i] = cc; S.s_frac[i] = frac; S.s_K[i] = K;
        });
        ex.sync();
        const int nblk = (m + FB200_BLK - 1) / FB200_BLK;  // blocks of unknowns first+1+8b .. first+8b+8
        const int nred = 2 * nblk;

        ex.tick(0);
        int iters = 0;
        bool exceeds;  // max_dif_rel(K1, K0) > eps: the loop needs the comparison only, so the CTA reduces the predicate, not the maximum
        do {
            // -- block elimination: every block thread reduces its <= 8 rows to two rows in (x_0, x_last) --
            ex.blocks(nblk, [&](int b) {
                const int i0 = first + 1 + b * FB200_BLK;          // first row of the block
                const int nb = (m - b * FB200_BLK < FB200_BLK) ? m - b * FB200_BLK : FB200_BLK;
                const int p0 = pidx(i0);
                double L0, D0, U0, R0, L1, D1, U1, R1;
                const double l0 = S.s_l[p0], d0 = S.s_d[p0], u0 = S.s_u[p0], r0 = S.s_rhs[p0];
                if (nb == 1) {
                    L0 = l0; D0 = d0; U0 = 0.; R0 = r0;   // a lone row can only be the last block: no upper coupling
                    L1 = 0.; D1 = 1.; U1 = 0.; R1 = 0.;   // dummy row, decoupled
                } else {
                    // forward: row j becomes  a_j x_0 + x_j + c_j x_{j+1} = g_j ; rolling, results parked in p_a/p_c/p_g
                    double a, c, g;
                    {
                        const int pj = pidx(i0 + 1);
                        const double inv = fb_rcp(S.s_d[pj]);
                        a = S.s_l[pj] * inv; c = S.s_u[pj] * inv; g = S.s_rhs[pj] * inv;
                        S.p_a[pj] = a; S.p_c[pj] = c; S.p_g[pj] = g;
                    }
                    // (A sweep over unnormalised rows, without the reciprocal on the critical path, is 0.4 % faster but five times
                    // less accurate on the Nx = 10^4 linear grid of the analytical tests: not used.)
                    for (int j = 2; j < nb; ++j) {
                        const int pj = pidx(i0 + j);
                        const double l = S.s_l[pj];
                        const double inv = fb_rcp(S.s_d[pj] - l * c);
                        a = -l * a * inv; c = S.s_u[pj] * inv; g = (S.s_rhs[pj] - l * g) * inv;
                        S.p_a[pj] = a; S.p_c[pj] = c; S.p_g[pj] = g;
                    }
                    L1 = a; D1 = 1.; U1 = c; R1 = g;      // row nb-1:  a x_0 + x_last + c x_next = g
                    // backward: row j becomes  a_j x_0 + x_j + c_j x_last = g_j   (1 <= j <= nb-2)
                    if (nb == 2) {
                        L0 = l0; D0 = d0; U0 = u0; R0 = r0;
                    } else {
                        int pj = pidx(i0 + nb - 2);
                        double an = S.p_a[pj], cn = S.p_c[pj], gn = S.p_g[pj];  // row nb-2 already couples to x_last
                        for (int j = nb - 3; j >= 1; --j) {
                            pj = pidx(i0 + j);
                            const double cj = S.p_c[pj];
                            an = S.p_a[pj] - cj * an; gn = S.p_g[pj] - cj * gn; cn = -cj * cn;
                            S.p_a[pj] = an; S.p_c[pj] = cn; S.p_g[pj] = gn;
                        }
                        // row 0 with x_1 = g_1 - a_1 x_0 - c_1 x_last substituted
                        L0 = l0; D0 = d0 - u0 * an; U0 = -u0 * cn; R0 = r0 - u0 * gn;
                    }
                }
                const double inv0 = fb_rcp(D0), inv1 = fb_rcp(D1);
                // reduced row q is stored at rpos(q): the identity, or the warp-major order of the hybrid solver below
                S.lu[0][epos(b)] = make_double2(L0 * inv0, U0 * inv0);
                S.rr[0][epos(b)] = R0 * inv0;
                S.lu[0][opos(b)] = make_double2(L1 * inv1, U1 * inv1);
                S.rr[0][opos(b)] = R1 * inv1;
            });
            ex.sync();
            ex.tick(8);
            // -- reduced system (nred unknowns): parallel cyclic reduction, diagonal normalised to 1 --
            int cur = 0;
#ifdef FB_PCR_ODD
            // (Variant, not the default: 8 % fewer executed instructions, the same time — the solve is bound by its chain of
            // dependent levels and barriers, not by issue slots.  126 GPU tests green with it.)
            // After the first level of a cyclic reduction the even and the odd unknowns form two independent systems, and a full
            // PCR solves both.  Only the odd one (the last unknown of every block, o_b) is solved here — level for level the
            // arithmetic a full PCR applies to its odd rows — and the even unknowns (the first of every block, e_b) follow from
            // their own level-0 rows:  e_b = r_b - l_b o_{b-1} - u_b o_b.  Half the rows at every level: half the warps.
            if constexpr (Ex::kWarpPcr) {
                // Hybrid over four warps: thread t < 128 owns o_p, p = warp + 4 lane.  The first level takes its three level-0
                // rows from buffer 0 (opos / epos); the strides 1, 2 (in p) couple different warps and go through shared memory;
                // after them the couplings are at distance 4 = the neighbouring LANE, and strides 4 .. 64 are warp shuffles.
                // Five barriers per Picard iteration, all of them reached by every thread.
                constexpr int H = FB200_NRED_MAX / 2;
                const int t = ex.thread(), lane = t & 31;
                const int p = (t >> 5) + 4 * lane;
                const bool mine = t < H && p < nblk;
                double lo = 0., up = 0., r = 0.;
                if (t < H) {
                    if (mine) {
                        const double2 me = S.lu[0][t], a = S.lu[0][H + t];
                        const double rme = S.rr[0][t], ra = S.rr[0][H + t];
                        double2 b = make_double2(0., 0.);
                        double rb = 0.;
                        if (p + 1 < nblk) { b = S.lu[0][epos(p + 1)]; rb = S.rr[0][epos(p + 1)]; }
                        const double inv = fb_rcp(1.0 - me.x * a.y - me.y * b.x);
                        lo = -me.x * a.x * inv; up = -me.y * b.y * inv; r = (rme - me.x * ra - me.y * rb) * inv;
                    }
                    S.lu[1][t] = make_double2(lo, up);
                    S.rr[1][t] = r;
                }
                ex.sync();
                if (t < H) {
                    if (mine) {  // stride 1
                        double2 a = make_double2(0., 0.), b = make_double2(0., 0.);
                        double ra = 0., rb = 0.;
                        if (p - 1 >= 0) { a = S.lu[1][ppos(p - 1)]; ra = S.rr[1][ppos(p - 1)]; }
                        if (p + 1 < nblk) { b = S.lu[1][ppos(p + 1)]; rb = S.rr[1][ppos(p + 1)]; }
                        const double inv = fb_rcp(1.0 - lo * a.y - up * b.x);
                        const double nlo = -lo * a.x * inv, nup = -up * b.y * inv;
                        r = (r - lo * ra - up * rb) * inv;
                        lo = nlo; up = nup;
                    }
                    // the odd level-0 rows in the lower half of buffer 0 are dead (read before the barrier above)
                    S.lu[0][t] = make_double2(lo, up);
                    S.rr[0][t] = r;
                }
                ex.sync();
                ex.tick(9);
                if (t < H) {
                    if (mine) {  // stride 2; the result stays in registers
                        double2 a = make_double2(0., 0.), b = make_double2(0., 0.);
                        double ra = 0., rb = 0.;
                        if (p - 2 >= 0) { a = S.lu[0][ppos(p - 2)]; ra = S.rr[0][ppos(p - 2)]; }
                        if (p + 2 < nblk) { b = S.lu[0][ppos(p + 2)]; rb = S.rr[0][ppos(p + 2)]; }
                        const double inv = fb_rcp(1.0 - lo * a.y - up * b.x);
                        const double nlo = -lo * a.x * inv, nup = -up * b.y * inv;
                        r = (r - lo * ra - up * rb) * inv;
                        lo = nlo; up = nup;
                    }
                    // No selects on the neighbours' values in the shuffle levels: a row without a lower (upper) neighbour at the
                    // current stride has lo (up) == +-0 exactly — the guards of the levels above made it a product with 0, and
                    // every later level multiplies it on — so whatever finite value the shuffle delivers (the lane's own one at the
                    // warp's edge, the zeros of a padding row) drops out as 0 * finite.  Padding rows (p >= nblk) hold 0 / 0 / 0.
                    for (int d = 1; d < 32; d <<= 1) {
                        const double a_x = ex.shfl_up(lo, d), a_y = ex.shfl_up(up, d), ra = ex.shfl_up(r, d);
                        const double b_x = ex.shfl_down(lo, d), b_y = ex.shfl_down(up, d), rb = ex.shfl_down(r, d);
                        const double inv = fb_rcp(1.0 - lo * a_y - up * b_x);
                        const double nlo = -lo * a_x * inv, nup = -up * b_y * inv;
                        r = (r - lo * ra - up * rb) * inv;
                        lo = nlo; up = nup;
                    }
                    // o_p into the final layout z[2 p + 1] (buffer 1: last read before the previous barrier)
                    if (mine) S.rr[1][2 * p + 1] = r;
                }
                ex.sync();
                if (mine) {  // e_p from its level-0 row (upper half of buffer 0, untouched since the block elimination)
                    const double2 e = S.lu[0][H + t];
                    const double om = (p > 0) ? S.rr[1][2 * p - 1] : 0.;
                    S.rr[1][2 * p] = S.rr[0][H + t] - e.x * om - e.y * r;
                }
                ex.sync();
                cur = 1;
            } else {
                constexpr int H = FB200_NRED_MAX / 2;
                // first level, odd rows only: row 2 p + 1 with its neighbours 2 p and 2 p + 2, into the lower half of buffer 1
                ex.reduced(nblk, [&](int p) {
                    const int q = 2 * p + 1;
                    const double2 me = S.lu[0][q], a = S.lu[0][q - 1];
                    const double rme = S.rr[0][q], ra = S.rr[0][q - 1];
                    double2 b = make_double2(0., 0.);
                    double rb = 0.;
                    if (q + 1 < nred) { b = S.lu[0][q + 1]; rb = S.rr[0][q + 1]; }
                    const double inv = fb_rcp(1.0 - me.x * a.y - me.y * b.x);
                    S.lu[1][p] = make_double2(-me.x * a.x * inv, -me.y * b.y * inv);
                    S.rr[1][p] = (rme - me.x * ra - me.y * rb) * inv;
                });
                ex.sync();
                // PCR of the nblk odd unknowns, ping-pong between the two halves of buffer 1 (buffer 0 keeps the even rows)
                int off = 0;
                for (int s = 1; s < nblk; s <<= 1) {
                    ex.reduced(nblk, [&](int p) {
                        const double2 me = S.lu[1][off + p];
                        const double rme = S.rr[1][off + p];
                        const int pm = p - s, pp = p + s;
                        double2 a = make_double2(0., 0.), b = make_double2(0., 0.);
                        double ra = 0., rb = 0.;
                        if (pm >= 0) { a = S.lu[1][off + pm]; ra = S.rr[1][off + pm]; }
                        if (pp < nblk) { b = S.lu[1][off + pp]; rb = S.rr[1][off + pp]; }
                        const double inv = fb_rcp(1.0 - me.x * a.y - me.y * b.x);
                        S.lu[1][(off ^ H) + p] = make_double2(-me.x * a.x * inv, -me.y * b.y * inv);
                        S.rr[1][(off ^ H) + p] = (rme - me.x * ra - me.y * rb) * inv;
                    });
                    ex.sync();
                    off ^= H;
                }
                // z[2 p + 1] = o_p, z[2 p] = e_p, in buffer 0 (a thread overwrites the two level-0 right-hand sides of its own block only)
                ex.reduced(nblk, [&](int p) {
                    const double op = S.rr[1][off + p];
                    const double om = (p > 0) ? S.rr[1][off + p - 1] : 0.;
                    const double2 e = S.lu[0][2 * p];
                    S.rr[0][2 * p] = S.rr[0][2 * p] - e.x * om - e.y * op;
                    S.rr[0][2 * p + 1] = op;
                });
                ex.sync();
                cur = 0;
            }
#else
            if constexpr (Ex::kWarpPcr) {
                // Hybrid: thread t owns unknown q = warp + 8 lane.  The strides 1, 2, 4 couple different warps and go
                // through shared memory (two barriers; the third level stays in registers); after them the couplings are
                // at distance 8 = the neighbouring LANE, and strides 8 .. 128 are warp shuffles without any barrier.
                // Same arithmetic, level by level, as the plain loop below (levels beyond nred see no neighbour and
                // leave a row unchanged bit for bit): 4 CTA barriers per Picard iteration instead of 9.
                const int t = ex.thread(), lane = t & 31;
                const int q = (t >> 5) + 8 * lane;
                double lo = 0., up = 0., r = 0.;
                for (int s = 1; s <= 4; s <<= 1) {
                    const double2 me = S.lu[cur][t];
                    const double rme = S.rr[cur][t];
                    const int qm = q - s, qp = q + s;
                    double2 a = make_double2(0., 0.), b = make_double2(0., 0.);
                    double ra = 0., rb = 0.;
                    if (q < nred && qm >= 0) { a = S.lu[cur][rpos(qm)]; ra = S.rr[cur][rpos(qm)]; }
                    if (qp < nred) { b = S.lu[cur][rpos(qp)]; rb = S.rr[cur][rpos(qp)]; }
                    const double inv = fb_rcp(1.0 - me.x * a.y - me.y * b.x);
                    lo = -me.x * a.x * inv; up = -me.y * b.y * inv; r = (rme - me.x * ra - me.y * rb) * inv;
                    if (s < 4) {
                        S.lu[cur ^ 1][t] = make_double2(lo, up);
                        S.rr[cur ^ 1][t] = r;
                        ex.sync();
                        cur ^= 1;
                    }
                }
                ex.tick(9);
                // No selects on the neighbours' values in the shuffle levels: a row without a lower (upper) neighbour at the
                // current stride has lo (up) == +-0 exactly — the guards of the levels above made it a product with 0, and
                // every later level multiplies it on — so whatever finite value the shuffle delivers (the lane's own one at the
                // warp's edge, the zeros of a padding row) drops out as 0 * finite.
                if (q >= nred) { lo = 0.; up = 0.; r = 0.; }  // padding rows: finite, and they stay 0 / 0 / 0
                for (int d = 1; d < 32; d <<= 1) {
                    const double a_x = ex.shfl_up(lo, d), a_y = ex.shfl_up(up, d), ra = ex.shfl_up(r, d);
                    const double b_x = ex.shfl_down(lo, d), b_y = ex.shfl_down(up, d), rb = ex.shfl_down(r, d);
                    const double inv = fb_rcp(1.0 - lo * a_y - up * b_x);
                    const double nlo = -lo * a_x * inv, nup = -up * b_y * inv;
                    r = (r - lo * ra - up * rb) * inv;
                    lo = nlo; up = nup;
                }
                // buffer 1 was last read two barriers ago; buffer 0 may still be read by a warp at stride 4
                if (q < nred) S.rr[1][q] = r;
                ex.sync();
                cur = 1;
            } else {
                for (int s = 1; s < nred; s <<= 1) {
                    ex.reduced(nred, [&](int q) {
                        const double2 me = S.lu[cur][q];
                        const double rme = S.rr[cur][q];
                        const int qm = q - s, qp = q + s;
                        double2 a = make_double2(0., 0.), b = make_double2(0., 0.);
                        double ra = 0., rb = 0.;
                        if (qm >= 0) { a = S.lu[cur][qm]; ra = S.rr[cur][qm]; }
                        if (qp < nred) { b = S.lu[cur][qp]; rb = S.rr[cur][qp]; }
                        const double inv = fb_rcp(1.0 - me.x * a.y - me.y * b.x);
                        S.lu[cur ^ 1][q] = make_double2(-me.x * a.x * inv, -me.y * b.y * inv);
                        S.rr[cur ^ 1][q] = (rme - me.x * ra - me.y * rb) * inv;
                    });
                    ex.sync();
                    cur ^= 1;
                }
            }
#endif
            const auto z = S.rr[cur];
            const bool incremental = iters > 0 || !has_C;
            ex.tick(1);
            // -- back-substitution fused with the y, W, K update and the convergence norm (nonlinear_diffusion.cpp:83-88) --
            // K = frac W / y = (frac h^n / ((1-m) D)) |y|^(1-m) / y  is a pure power of y, so from the second iteration on
            // (the first K holds the wind term tau C, nonlinear_diffusion.cpp:61 vs :86)  K1 = K0 (y_old / y)^m.
            // With x = y_old / y - 1 small, (1 + x)^m - 1 is an 8-term binomial series (|x| < 2^-6: truncation < 1e-18),
            // 10 FP64 operations instead of the 60 of a pow; |K1 - K0| / |K1| = |s| / (1 + s) follows from the same s.
            if (incremental) {
                // Pass 1 is branch-free (loads at the thread's own, always valid indices; selects and predicated stores), so the
                // chains of a thread's four rings — reciprocal, series — interleave instead of running one after the other; the
                // rings whose change is too large for the series are marked (NaN diagonal) and redone with the pow in pass 2.
                double vmax = 0.;
                bool anybad = false;
                const double b0 = M.m_binom[0], b1 = M.m_binom[1], b2 = M.m_binom[2], b3 = M.m_binom[3], b4 = M.m_binom[4],
                             b5 = M.m_binom[5], b6 = M.m_binom[6], b7 = M.m_binom[7];
                ex.points_ilp([&](int i, int k) {
                    const bool in = (i > first) && (i <= last);
                    const int un = i - first - 1;
                    int b = ((un < 0) ? 0 : un) / FB200_BLK;
                    b = (b > nblk - 1) ? nblk - 1 : b;
                    const int j = un - b * FB200_BLK;
                    const int nb = (m - b * FB200_BLK < FB200_BLK) ? m - b * FB200_BLK : FB200_BLK;
                    const int pi = pidx(i);
                    const double z0 = z[2 * b], z1 = z[2 * b + 1];
                    const double yg = S.p_g[pi] - S.p_a[pi] * z0 - S.p_c[pi] * z1;
                    const double y = (j == 0) ? z0 : ((j == nb - 1) ? z1 : yg);
                    const double y_old = S.F[i];
                    const double K0 = S.s_K[i];
                    const double inv = fb_rcp14(y);  // x is a small correction: 1e-14 of it is 1e-16 of K
                    const double x = (y_old - y) * inv;
                    double sr = b7;
                    sr = fma(sr, x, b6); sr = fma(sr, x, b5); sr = fma(sr, x, b4); sr = fma(sr, x, b3);
                    sr = fma(sr, x, b2); sr = fma(sr, x, b1); sr = fma(sr, x, b0);
                    sr *= x;
                    const double K1 = fma(K0, sr, K0);
                    const double rel = fabs(sr) * (1. - sr);  // |s| / (1 + s) to O(s^2); only compared with eps
                    const bool upd = in && (i != last);       // K[last] stays at its pre-step value
                    const bool ok = fabs(x) < 0.015625;
                    if (i == first) S.F[i] = F_in;
                    if (in) S.F[i] = y;
                    if (upd) {
                        S.s_K[i] = ok ? K1 : K0;
                        S.s_d[pi] = ok ? S.s_cc[i] + K1 : NAN;
                    }
                    vmax = fmax(vmax, (upd && ok) ? rel : 0.);  // fmax skips a NaN, as `x > max` does in max_dif_rel
                    anybad = anybad || (upd && !ok);
                });
                if (anybad) {
                    int nbad = 0;
                    ex.points([&](int i, int k) {
                        if (!(i > first && i < last)) return;
                        const int pi = pidx(i);
                        if (!fb_isnan(S.s_d[pi])) return;
                        ++nbad;
                        const double y = S.F[i], K0 = S.s_K[i];
                        const double W = wunc_point(M, y, S.hn[i]);
                        const double K1 = S.s_frac[i] * W * fb_rcp(y);
                        S.s_K[i] = K1;
                        S.s_d[pi] = S.s_cc[i] + K1;
                        vmax = fmax(vmax, fabs((K1 - K0) * fb_rcp(fabs(K1))));
                    });
                    ex.note_pow_fallbacks(nbad);
                }
                exceeds = ex.any_of(vmax > M.eps);  // vmax is never NaN (fmax skipped them, as `x > max` does in max_dif_rel)
            } else {
                ex.count(FB_CNT_POW_EVALS, m - 1);
                const double maxrel = ex.reduce_max([&](int i, int k) -> double {
                    if (i == first) S.F[i] = F_in;
                    if (!(i > first && i <= last)) return 0.;
                    const int un = i - first - 1, b = un / FB200_BLK, j = un - b * FB200_BLK;
                    const int nb = (m - b * FB200_BLK < FB200_BLK) ? m - b * FB200_BLK : FB200_BLK;
                    const int pi = pidx(i);
                    double y;
                    if (j == 0) y = z[2 * b];
                    else if (j == nb - 1) y = z[2 * b + 1];
                    else y = S.p_g[pi] - S.p_a[pi] * z[2 * b] - S.p_c[pi] * z[2 * b + 1];
                    S.F[i] = y;
                    if (i == last) return 0.;  // K[last] stays at its pre-step value
                    const double K0 = S.s_K[i];
                    const double W = wunc_point(M, y, S.hn[i]);
                    const double K1 = S.s_frac[i] * W * fb_rcp(y);  // d = c0 + K absorbs far more than the last ulp of K
                    S.s_K[i] = K1;
                    S.s_d[pi] = S.s_cc[i] + K1;
                    return fabs((K1 - K0) * fb_rcp(fabs(K1)));  // only compared with eps; a NaN is skipped by the reduction
                });
                exceeds = maxrel > M.eps;
            }
            ex.tick(2);
            ++iters;
            if (iters >= 100000) return -1;
        } while (exceeds);
        ex.sync();
        k_valid = true;
        ex.count(FB_CNT_RING_STEPS, m);
        ex.count(FB_CNT_RING_ITERS, (long long)m * iters);
        return iters;
    }

    // ---- structure pass + truncateOuterRadius + one row ----
    // do_truncate: after a step (truncateOuterRadius, freddi_evolution.cpp:32-74); false for row 0.
    // Returns false on RadiusCollapse (no row is written then).  row may be null (no output wanted).
    FB_HD bool structure_and_row(bool do_truncate, double* row, double* Fsnap_row, int* bounds_row) {
        const int Nx = M.Nx;
        const int first = st.first;
        int last = st.last;
        const StateScalars sc = state_scalars(ex, M, S, first, true);

        // hot-zone structure of every point of [first, last]
        // After a solve the closure value of an interior ring needs no pow: the last Picard iteration left
        // K_i = frac_i W(y_i) / y_i for the final y (nonlinear_diffusion.cpp:85-87), so W_i = K_i y_i / frac_i (K[last] is
        // frozen at its pre-step value and `first` is not part of the system: those two use the closure itself).
        // Height is needed at every ring only if H/R enters Qx (plane emitters, irrindex != 0) or the star's shadow.
        const bool height_everywhere = (M.irrindex != 0.) || M.angdist_disk != FB200_ANGDIST_ISOTROPIC ||
                                       (M.is_ns && M.angdist_ns != FB200_ANGDIST_ISOTROPIC) || (kExtras && cfg.star_flux);
        ex.points([&](int i, int k) {
            if (i >= Nx) return;
            double Tph = 0., Tirr = 0., Sigma = 0., Tvis = 0.;
            if (i >= first && i <= last) {
                if (i == last) {  // the outer edge feeds six columns and the truncation test: the reference's own operation order
                    const PointStructure q = hot_structure(M, sc, S.h[i], S.R[i], S.hn[i], S.hm175[i], S.hotR[i], S.F[i]);
                    Tph = q.Tph; Tirr = q.Tirr; Sigma = q.Sigma; Tvis = q.Tvis;
                    ex.scalar(0) = q.Kirr; ex.scalar(1) = q.Height / S.R[i];
                    if (kExtras && cfg.star_flux) S.aux[i] = q.Height / S.R[i];
                } else {
                    const double Wk = (do_truncate && i > first) ? S.s_K[i] * S.F[i] * fb_rcp(S.s_frac[i]) : -1.;
                    const RowStructure q = hot_structure_lean(M, sc, S.R[i], S.hn[i], S.sgA[i], S.qB[i], S.tv4[i], S.hotR[i], S.F[i], Wk, height_everywhere);
                    Tph = q.Tph; Tirr = q.Tirr; Sigma = q.Sigma;
                    if (kExtras && cfg.star_flux) S.aux[i] = q.H2R;  // for Height2R of the star's irradiation sources
                }
            }
            S.Tph[i] = Tph; S.Tirr[i] = Tirr; S.Sigma[i] = Sigma; S.Tvis[i] = Tvis;
        });
        ex.sync();

        if (do_truncate && M.Thot > 0. && fb_isfinite(st.Mdot_in_prev) && !(sc.Mdot < 0.0) && !(st.Mdot_in_prev < 0.0) &&
            !(sc.Mdot > st.Mdot_in_prev)) {
            int branch;  // 0: cooling front, 1: Teff, 2: Tirr
            if (S.Tirr[last] / S.Tvis[last] < M.Tirr2Tvishot) branch = 0;
            else branch = (M.boundcond == FB200_BOUNDCOND_TEFF) ? 1 : 2;
            const double Sigma_last = S.Sigma[last], R_last = S.R[last];
            // the do-while walks ii = last, last-1, ... and stops at the first ii whose condition is false
            const int jmax = ex.reduce_max_int([&](int i, int k) -> int {
                if (!(i > first && i <= last)) return -1;
                bool cond;
                if (branch == 0) {
                    const double r = S.R[i];
                    const double R_cf = R_last - v_cooling_front(M, r, Sigma_last) * M.tau;
                    cond = (r > R_cf) && (S.Sigma[i] < sigma_minus(M, r)) && (S.Tirr[i] < M.Thot);
                } else if (branch == 1) {
                    cond = S.Tph[i] < M.Thot;
                } else {
                    cond = S.Tirr[i] < M.Thot;
                }
                return cond ? -1 : i;
            });
            if (jmax < 0) return false;  // ii <= first: RadiusCollapseException
            if (jmax <= last - 1) {
                last = jmax;
                st.last = last;
                ex.points([&](int i, int k) {  // the columns of the new outer edge
                    if (i == last) {
                        const PointStructure q = hot_structure(M, sc, S.h[i], S.R[i], S.hn[i], S.hm175[i], S.hotR[i], S.F[i]);
                        ex.scalar(0) = q.Kirr; ex.scalar(1) = q.Height / S.R[i];
                        S.Tph[i] = q.Tph; S.Tirr[i] = q.Tirr; S.Sigma[i] = q.Sigma; S.Tvis[i] = q.Tvis;
                    }
                });
                ex.sync();
            }
        }
        ex.tick(3);
        if (!row) return true;
        // BasicFreddiIrradiationSource::Height2R (cpp/src/freddi_state.cpp:663-671): max of Height/R over [first, Nx); the cold
        // zone has Height = h2rcold R (freddi_state.cpp:271-273).  Taken here, before S.aux is reused below.
        double star_h2r = 0.;
        if (kExtras && cfg.star_flux) {
            star_h2r = ex.reduce_max([&](int i, int k) -> double {
                if (i < first || i >= Nx) return 0.;
                if (i <= last) return S.aux[i];
                return (M.h2rcold * S.R[i]) / S.R[i];
            });
            if (!(star_h2r > 0.)) star_h2r = 0.;  // std::max(H/R, 0.0) chain: NaN and negatives leave 0
            ex.sync();
            // before the radial sums below: star_row reuses the reduction slots sum(0..2)
            star_row(sc, star_h2r, do_truncate);  // the state at t0 has no sources yet (freddi_state.cpp:79, freddi_evolution.cpp:28)
        }

        // ---- Tph_X over the hot zone, freddi_state.cpp:292-312 ----
        const double absMdot = fabs(sc.Mdot_plain);
        const double KM = 3. * absMdot * p6(FB_C_LIGHT) / (8. * FB_PI * p2(M.GM));
        const double F_first = S.F[first];
        ex.points([&](int i, int k) {
            if (i >= Nx) return;
            double T = 0., wq = 0.;
            if (i >= first && i <= last) {
                double x = S.c1[i] * F_first;
                const double g = KM * S.Gb[i];  // = T_GR^4 sigma; T_GR = pow(., 0.25) is NaN for a negative argument
                x += (g < 0.) ? FB_NAN_INVALID : g;
                T = M.colourfactor * pow025(x * (1. / FB_SIGMA_SB));
                // weight of this ring in the radial trapezoid of the hot zone: 2 pi R_i (R_{i+1} - R_{i-1}) (ends one-sided)
#ifdef FB_OLD_WQ
                if (first < last) {
                    const double R = S.R[i];
                    double w;
                    if (i == first) w = S.R[i + 1] - R;
                    else if (i == last) w = R - S.R[i - 1];
                    else w = S.R[i + 1] - S.R[i - 1];
                    wq = 2 * FB_PI * R * w;
                }
#else
                if (first < last) wq = 2 * FB_PI * S.R[i] * trapz_weight(S.R, i, first, last);
#endif
            }
            S.TphX[i] = T;
            S.wq[i] = wq;
        });
        ex.sync();
        ex.tick(10);

        // max Tph_X over [first, last] with std::max_element's NaN behaviour (output.cpp:208)
        const double TphX_hottest = ex.reduce_max([&](int i, int k) -> double {
            if (i >= first && i <= last) return S.TphX[i];
            return -INFINITY;
        });
        double TphXmax = TphX_hottest;
        if (fb_isnan(S.TphX[first])) TphXmax = S.TphX[first];  // max_element returns the first element
        // Rings that cannot contribute to Lx are not integrated (the reference walks them all).  With a = h nu1 / k T,
        // B_nu(T_i) / B_nu(T_hot) <= exp(-(a_i - a_hot)) for every nu of the band, and Lx >= the hottest ring's share
        // 1/2 wq_hot I(T_hot), so the rings with a_i - a_hot > P change Lx by less than
        //     exp(-P) * sum_i wq_i / wq_hot  <=  exp(-P) * 4 pi R_last^2 / wq[first]
        // relative (sum_i wq_i = 2 pi sum R_i (R_{i+1} - R_{i-1}) <= 4 pi R_last^2; wq[first], the one-sided weight of the
        // innermost ring, is the smallest weight of the hot zone on the log and the linear grid alike).  P is chosen per row so
        // that this bound is 1e-16 — six orders below the 1e-10 parity tolerance, below the rounding of the sum itself — and
        // never above the fixed FB200_LX_PRUNE = 100.
        const double a_hot = (TphX_hottest > 0.) ? FB_H_OVER_KB * M.emin / TphX_hottest : 0.;
        double prune = FB200_LX_PRUNE;
        {
            const double ratio = 4. * FB_PI * p2(S.R[last]) * fb_rcp(S.wq[first]);
            if (ratio > 1. && ratio < 1e25) {
                const double P = log(ratio) + 36.85;  // -ln(1e-16) = 36.84
                prune = (P < FB200_LX_PRUNE) ? P : FB200_LX_PRUNE;
            }
        }
        const double T_prune = (cfg.lx_prune && TphX_hottest > 0.) ? FB_H_OVER_KB * M.emin / (a_hot + prune) : 0.;

        // ---- radial trapezoids (util.cpp:9-41): Mdisk, Lx, Fnu[k] (hot), Fnu[k] (cold) ----
        // slots: 0 Mdisk, 1 Lx, [2, 2+nl) Fnu hot, [2+nl, 2+2nl) Fnu cold, [q_pb, q_pb+npb) passbands hot,
        //        [q_pb+npb, q_pb+2npb) passbands cold
        const int nl = cfg.n_lambdas, npb = kExtras ? cfg.n_passbands : 0;
        const int per = cfg.cold_disk ? 2 : 1;
        const int q_pb = 2 + nl * per;
        const int nq = q_pb + npb * per;
        double cold_K = 0.;
        const double cold_mu = M.h2rcold;  // Height/R of the cold zone, freddi_state.cpp:271-273
        if (cfg.cold_disk) cold_K = M.Cirr_cold * ((M.irrindex_cold == 0.) ? 1.0 : pow((M.h2rcold * 1.0) / (1.0 * 0.05), M.irrindex_cold));
        const double* lx_tab = M.lx_tab ? cfg.lx_table : nullptr;
        int my_trials = 0;  // try_step calls of this thread's rings (roofline bookkeeping): summed by ex.note_lx_trials, not by a reduction pass
        ex.reduce_sums(nq, [&](int i, int k, int q) -> double {
            if (i >= Nx) return 0.;
            const bool hot_pb = (q >= q_pb && q < q_pb + npb);
            if (q < 2 + nl || hot_pb) {  // hot zone [first, last]: weights prepared above (0 outside)
                const double wq = S.wq[i];
                if (wq == 0.) return 0.;
                double y;
                if (hot_pb) {
                    const int p = q - q_pb;
                    y = passband_ring_sum(S.Tph[i], cfg.pb_a + cfg.pb_off[p], cfg.pb_w + cfg.pb_off[p], cfg.pb_off[p + 1] - cfg.pb_off[p]);
                } else if (q == 0) {
                    y = S.Sigma[i];
                } else if (q == 1) {
                    if (!cfg.compute_lx) return 0.;
                    int trials = 0;
                    const double T = S.TphX[i];
                    y = (T < T_prune) ? 0. : planck_nu1_nu2_tab(lx_tab, T, M.emin, M.emax, FB_KL(CK_TOL), &trials);
                    my_trials += trials;
                } else {
                    y = planck_lambda_pref(S.Tph[i], cfg.lambdas[q - 2], ex.lambda_pref(q - 2));
                }
                return y * wq;
            }
            // cold zone [last + 1, Nx - 1] (--colddiskflux): Height = h2rcold R, Kirr from the cold parameters
            const int rf = last + 1, rl = Nx - 1;
            if (rf >= rl || i < rf || i > rl) return 0.;
            const double R = S.R[i];
            const double w = trapz_weight(S.R, i, rf, rl);
            double Qx;
            if (M.is_ns)
                Qx = cold_K * (sc.Lbol_disk * angular_dist(M.angdist_disk, cold_mu) + sc.Lbol_ns * angular_dist(M.angdist_ns, cold_mu)) /
                     (4. * FB_PI * p2(R));
            else
                Qx = cold_K * sc.Lbol_disk * angular_dist(M.angdist_disk, cold_mu) / (4. * FB_PI * p2(R));
            const double Tcold = pow025(Qx / FB_SIGMA_SB);  // Tirr of the cold zone, freddi_state.cpp:221-231
            double y;
            if (q >= q_pb) {
                const int p = q - q_pb - npb;
                y = passband_ring_sum(Tcold, cfg.pb_a + cfg.pb_off[p], cfg.pb_w + cfg.pb_off[p], cfg.pb_off[p + 1] - cfg.pb_off[p]);
            } else {
                const int kl = q - 2 - nl;
                y = planck_lambda_pref(Tcold, cfg.lambdas[kl], ex.lambda_pref(kl));
            }
            return 2 * FB_PI * R * y * w;
        });
        ex.tick(4);
        // ---- the row (output.cpp:197-233, ns_output.cpp:6-19) ----
        ex.note_lx_trials(my_trials);
        ex.count(FB_CNT_ROW_RINGS, last - first + 1);
        const double Mdisk = 0.5 * ex.sum(0);
        const double Lx = cfg.compute_lx ? (2. * FB_PI * (0.5 * ex.sum(1))) / p4(M.colourfactor) : NAN;
        const double d2 = 4.0 * FB_PI * p2(M.distance);
        const double psi = angular_dist(M.angdist_disk, M.cosi);
        const int ncol = cfg.n_cols;
        const double t_now = st.t;
        // One thread per column, columns dealt round-robin over 8 warps (column c -> lane c / 8 of warp c % 8): every case of the
        // switch below is its own divergent path, and 18 of them in one warp cost more than the rest of the row phase.
        ex.points([&](int ii, int k) {
#if FB200_NX_MAX >= 256
            if (ii >= 256) return;
            const int i = (ii & 31) * 8 + (ii >> 5);
#else
            const int i = ii;  // small-grid builds (one warp per model): a point index per column
#endif
            if (i >= ncol) return;
            double v;
            switch (i) {
                case FB200_COL_T: v = t_now / 86400.; break;
                case FB200_COL_MDOT: v = sc.Mdot; break;
                case FB200_COL_MDISK: v = Mdisk; break;
                case FB200_COL_RHOT: v = S.R[last] / FB_SOLAR_RADIUS; break;
                case FB200_COL_SIGMAOUT: v = S.Sigma[last]; break;
                case FB200_COL_KIRROUT: v = ex.scalar(0); break;
                case FB200_COL_H2R: v = ex.scalar(1); break;
                case FB200_COL_TEFFOUT: v = S.Tph[last]; break;
                case FB200_COL_TIRROUT: v = S.Tirr[last]; break;
                case FB200_COL_QIRR2QVISOUT: v = p4(S.Tirr[last] / S.Tvis[last]); break;
                case FB200_COL_TPHXMAX: v = TphXmax * FB_K_BOLTZ / (1000. * FB_ELECTRON_VOLT); break;
                case FB200_COL_LX: v = Lx; break;
                case FB200_COL_LBOL: v = sc.Lbol_disk; break;
                case FB200_COL_FX: v = Lx * psi / d2; break;
                case FB200_COL_FBOL: v = sc.Lbol_disk * psi / d2; break;
                default: {
                    // per wavelength / passband: hot, [cold], [star, star_min, star_max]  (output.cpp:219-291)
                    const int k = i - FB200_N_BH_COLS;
                    const int percol = per + ((kExtras && cfg.star_flux) ? 3 : 0);
                    if (k < (nl + npb) * percol) {
                        const int kl = k / percol;          // wavelength kl < nl, else passband kl - nl
                        const int sub = k % percol;
                        if (sub >= per) {
                            v = ex.star_sum(3 * kl + (sub - per));
                        } else if (kl < nl) {
                            const bool cold = (sub == 1);
                            const double lam = cfg.lambdas[kl];
                            const double I = 0.5 * ex.sum(cold ? 2 + nl + kl : 2 + kl);
                            v = I * p2(lam) / FB_C_LIGHT * M.cosiOverD2;  // freddi_state.hpp:405-407
                        } else {
                            const int kp = kl - nl;
                            const bool cold = (sub == 1);
                            // the two nested trapezoids (radius, wavelength) each carry a factor 1/2
                            const double intens = 0.5 * (0.5 * ex.sum(cold ? q_pb + npb + kp : q_pb + kp));
                            v = intens * M.cosiOverD2 / cfg.pb_tdnu[kp];  // freddi_state.hpp:408-417
                        }
                    } else {
                        v = ns_column(k - (nl + npb) * percol, sc);
                    }
                }
            }
            row[i] = v;
        });
        if (Fsnap_row) {
            ex.points([&](int i, int k) { if (i < Nx) Fsnap_row[i] = S.F[i]; });
        }
        if (bounds_row) {
            ex.points([&](int i, int k) { if (i == 0) { bounds_row[0] = first; bounds_row[1] = last; } });
        }
        ex.sync();
        ex.tick(5);
        return true;
    }

    // ---- companion star (--starflux) ----
    // IrradiatedStar::Qirr / Star::Teff (cpp/src/star.cpp:68-80,246-258) with the sources FreddiState::star_irr_sources sets
    // at the end of every step (cpp/src/freddi_state.cpp:348-352,656-690; NS adds its own, ns_evolution.cpp:480-484), then
    // flux_star(lambda | passband, phase) for the current phase and the two conjunctions (freddi_state.cpp:355-361,
    // star.cpp:100-111, output.cpp:234-255,272-291).  Results in ex.star_sum(3 l + d).
    FB_HD void star_row(const StateScalars& sc, double h2r_max, bool with_sources) {
        star_teff(sc, h2r_max, with_sources);
        const double phase = 2.0 * FB_PI * (st.t - M.star_t0) / M.star_period;  // phase_opt, freddi_state.cpp:166-168
        const int nl = cfg.n_lambdas, npb = cfg.n_passbands;
        const double over_d2 = 4. * FB_PI * p2(M.distance);
        for (int l = 0; l < nl + npb; ++l) {
            if (l < nl) star_dir_sums(cfg.lambdas[l], ex.lambda_pref(l), -1, phase);
            else star_dir_sums(0., 0., l - nl, phase);
            ex.points([&](int i, int k) {
                if (i < 3) ex.star_sum(3 * l + i) = ex.sum(i) * (4. * FB_PI) / over_d2;
            });
            ex.sync();
        }
    }

    // Effective temperature of every triangle -> ex.star_T (IrradiatedStar::Qirr, Star::Teff)
    FB_HD void star_teff(const StateScalars& sc, double h2r_max, bool with_sources) {
        const int nt = M.star_ntri;
        const double* G = cfg.star_data + M.star_off;
        const double *cx = G, *cy = G + nt, *cz = G + 2 * nt, *nx = G + 3 * nt, *ny = G + 4 * nt, *nz = G + 5 * nt;
        const double h2r2 = p2(h2r_max);  // DiskShadowSource, star.cpp:190-196
        const double sx = -M.star_semiaxis, albedo = M.star_albedo;
        const double Tth4 = p4(M.star_Topt);
        const double L_disk = sc.Lbol_disk, L_ns = sc.Lbol_ns;
        const bool two = M.is_ns != 0;
        ex.tri_points(nt, [&](int t) {
            double Q = 0.;
            if (with_sources) {
                const double dx = cx[t] - sx, dy = cy[t] - 0.0, dz = cz[t] - 0.0;  // PointLikeSource::irr_flux, star.cpp:128-136
                const double len = sqrt(dx * dx + dy * dy + dz * dz);
                const double inv = 1.0 / len;
                const double ux = dx * inv, uy = dy * inv, uz = dz * inv;
                const double rho2 = p2(ux) + p2(uy);
                if (!((p2(uz) / rho2) < h2r2)) {
                    double cos_obj = -(ux * nx[t] + uy * ny[t] + uz * nz[t]);
                    if (cos_obj <= 0.0) cos_obj = 0.0;
                    const double cos_source = fabs(0.0 * ux + 0.0 * uy + 1.0 * uz);  // plain_normal = UnitVec3(0, 0)
                    const double lum_d = (M.angdist_disk == FB200_ANGDIST_ISOTROPIC) ? L_disk : L_disk * 2.0 * cos_source;
                    Q += lum_d * (1.0 - albedo) / (4. * FB_PI * p2(len)) * cos_obj;
                    if (two) {
                        const double lum_n = (M.angdist_ns == FB200_ANGDIST_ISOTROPIC) ? L_ns : L_ns * 2.0 * cos_source;
                        Q += lum_n * (1.0 - albedo) / (4. * FB_PI * p2(len)) * cos_obj;
                    }
                }
            }
            ex.star_T(t) = pow(Q / FB_SIGMA_SB + Tth4, 0.25);
        });
        ex.sync();
    }

    // Star::integrate(func, direction) (star.cpp:58-64) for the observer directions UnitVec3(inclination, phi) with
    // phi = phase, 0 and pi -> ex.sum(0..2).  func = B_lambda lambda^2 / c at wavelength lam (star.cpp:100-105), or
    // EnergyPassband::bb_nu of passband p >= 0 (star.cpp:107-111).
    FB_HD void star_dir_sums(double lam, double pref, int p, double phase) {
        const int nt = M.star_ntri;
        const double* G = cfg.star_data + M.star_off;
        const double *nx = G + 3 * nt, *ny = G + 4 * nt, *nz = G + 5 * nt, *ar = G + 6 * nt;
        const double si = M.star_sini, ci = M.star_cosi;
        const double d0x = si * cos(phase), d0y = si * sin(phase);
        const double d1x = si * 1.0, d1y = si * 0.0;
        const double d2x = si * -1.0, d2y = si * 1.2246467991473532e-16;  // cos(M_PI), sin(M_PI) in IEEE double
        ex.reduce_sums3(nt, [&](int t, double& a, double& b, double& c) {
            const double T = ex.star_T(t);
            double y;
            if (p < 0) {
                y = planck_lambda_pref(T, lam, pref) * p2(lam) / FB_C_LIGHT;
            } else {  // EnergyPassband::bb_nu, passband.hpp:43
                y = (0.5 * passband_ring_sum(T, cfg.pb_a + cfg.pb_off[p], cfg.pb_w + cfg.pb_off[p], cfg.pb_off[p + 1] - cfg.pb_off[p])) / cfg.pb_tdnu[p];
            }
            const double nxt = nx[t], nyt = ny[t], nzc = nz[t] * ci, area = ar[t];
            const double c0 = d0x * nxt + d0y * nyt + nzc, c1 = d1x * nxt + d1y * nyt + nzc, c2 = d2x * nxt + d2y * nyt + nzc;
            a = (c0 <= 0.) ? 0. : area * c0 * y;  // Triangle::area_cos, geometry.cpp:201-207
            b = (c1 <= 0.) ? 0. : area * c1 * y;
            c = (c2 <= 0.) ? 0. : area * c2 * y;
        });
    }

    // NS columns (ns_output.cpp:6-19); each evaluated by one thread
    FB_HD double ns_column(int k, const StateScalars& sc) const {
        const int first = st.first;
        const double area = 4 * FB_PI * M.hot_spot_area * p2(M.R_x);
        switch (k) {
            case FB200_NSCOL_RIN: return S.R[first];
            case FB200_NSCOL_ETANS: return sc.eta_ns;
            case FB200_NSCOL_LXNS:
            case FB200_NSCOL_FXNS: {  // ns_evolution.cpp:398-422
                const double T = pow025(sc.Lbol_ns_rest / (area * FB_SIGMA_SB));
                int trials;
                const double intensity = cfg.compute_lx ? planck_nu1_nu2_tab(M.lx_tab ? cfg.lx_table : nullptr, T, M.emin / M.redshift,
                                                                             M.emax / M.redshift, FB_KL(CK_TOL), &trials)
                                                        : NAN;
                const double Lx_ns = M.redshift * (area * FB_PI * intensity);
                if (k == FB200_NSCOL_LXNS) return Lx_ns;
                return Lx_ns * angular_dist(M.angdist_ns, M.cosi) / (4.0 * FB_PI * p2(M.distance));
            }
            case FB200_NSCOL_LBOLNS: return sc.Lbol_ns;
            case FB200_NSCOL_FBOLNS: return sc.Lbol_ns * angular_dist(M.angdist_ns, M.cosi) / (4.0 * FB_PI * p2(M.distance));
            case FB200_NSCOL_THOTSPOT: return pow025(sc.Lbol_ns_rest / (area * FB_SIGMA_SB)) * FB_K_BOLTZ / (1000. * FB_ELECTRON_VOLT);
            case FB200_NSCOL_FPIN: return sc.fp;
            case FB200_NSCOL_FMAGNIN: return ex.Fmagn(first);
            case FB200_NSCOL_FIN: return S.F[first];
            default: return NAN;
        }
    }

    // ---- one FreddiEvolution::step (freddi_evolution.cpp:17-29) followed by the row of the new state ----
    // Returns false when the model stops (collapse / failure); st.status says why.
    FB_HD bool step_and_row(double* row, double* Fsnap_row, int* bounds_row) {
        if (M.is_ns) {
            if (!truncate_inner()) { st.status = FB200_MODEL_COLLAPSED; return false; }
        }
        // FreddiState::step, freddi_state.cpp:147-153
        const StateScalars sc = state_scalars(ex, M, S, st.first, false);
        st.Mdot_in_prev = sc.Mdot;
        st.i_t += 1;
        st.t += M.tau;
        ex.tick(6);
        const int it = solve_step(sc.Mdot);
        if (it < 0) { st.status = FB200_MODEL_FAILED; return false; }
        if (!structure_and_row(true, row, Fsnap_row, bounds_row)) { st.status = FB200_MODEL_COLLAPSED; return false; }
        ex.count(FB_CNT_PICARD, it);  // counted like rows_valid: completed steps only
        return true;
    }
};

}  // namespace fb200
#endif
