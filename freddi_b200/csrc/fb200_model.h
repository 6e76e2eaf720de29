// Resolved per-model constants shared by the host resolver (resolve.cpp) and the kernels (evolve.cu).
//
// One fb200_model_dev_t is what the reference spreads over FreddiState::DiskStructure
// (cpp/include/freddi_state.hpp:218-240), OpacityRelated (cpp/include/opacity_related.hpp),
// the Arguments blocks (cpp/include/arguments.hpp:46-333) and NeutronStarStructure
// (cpp/include/ns/ns_evolution.hpp:42-67), reduced to the scalars the time step and the light-curve
// row read.  All values CGS.  Everything that needs libm on constants (cbrt, acos, cos, pow of
// per-model scalars) is evaluated once on the host so that discrete decisions the kernels make from
// them (e.g. `x < x0` in T_GR, cpp/src/spectrum.cpp:47-50) see the same bits as the reference.
#ifndef FB200_MODEL_H
#define FB200_MODEL_H

#include <stdint.h>

/* Radial points a kernel build handles.  The product library holds two builds of the same engine: the shared-memory
 * resident one (1024: every BASELINE configuration) and, compiled with -DFB_BIG, a large-grid one whose per-point
 * arrays live in an L2-resident global scratch slot (FB200_NX_LIMIT: the reference's analytical tests run Nx = 10^4,
 * python/test/test_analytical.py:84,113,216). */
#ifndef FB200_NX_MAX
#define FB200_NX_MAX 1024
#endif
#define FB200_NX_LIMIT 16384

/* physical constants: cpp/include/gsl_const_cgsm.h:24-38,107,113-114 (GSL 2009 CGSM values) */
#define FB_C_LIGHT 2.99792458e10
#define FB_G_GRAV 6.673e-8
#define FB_H_PLANCK 6.62606896e-27
#define FB_K_BOLTZ 1.3806504e-16
#define FB_SIGMA_SB 5.67040047374e-5
#define FB_M_PROTON 1.67262158e-24
#define FB_M_ELECTRON 9.10938188e-28
#define FB_THOMSON 6.65245893699e-25
#define FB_SOLAR_MASS 1.98892e33
#define FB_SOLAR_RADIUS 6.955e10 /* cpp/include/constants.hpp:8 */
#define FB_ELECTRON_VOLT 1.602176487e-12
#define FB_PI 3.14159265358979323846

typedef struct fb200_model_dev {
    /* --- grid / run --- */
    int32_t Nx, Nt, first0, is_ns;
    int32_t opacity, boundcond, windtype, angdist_disk;
    int32_t angdist_ns, fptype, kappattype, nsprop;
    int32_t wind_static;  /* 1: A/B/C arrays uploaded (static winds), 0: none */
    int32_t wind_dynamic; /* 1: C (or B) recomputed from Mdot_in every step */
    int32_t has_AB;       /* wind A or B can be non-zero (coefficients a,b,c0 then depend on the wind) */
    int32_t lx_tab;       /* 1: emax / emin is the band ratio of the ensemble's Lx table (OutCfg::lx_table, lx_table.h) */
    double tau, eps, t0;
    /* --- central object / geometry --- */
    double GM, Mx, kerr, eta, R_g, cosi, distance, cosiOverD2;
    double alpha, alphacold, mu;
    /* --- opacity closure, cpp/src/opacity_related.cpp:27-93 --- */
    double one_minus_m, n_op, D, Height_coef, Height_exp_F, Height_exp_R;
    double m_binom[8]; /* binomial coefficients C(m, k), k = 1..8: (1 + x)^m - 1 for the incremental K update of the Picard loop */
    /* --- boundary / truncation --- */
    double Mdot_out, F_in0, Thot, Tirr2Tvishot;
    /* Sigma_minus/Sigma_plus/v_cooling_front (cpp/src/freddi_state.cpp:692-715): per-model factors */
    double sigma_minus_k, sigma_plus_k; /* 74.6*pow(alphacold/0.1,-0.83) [*] pow(Mx/Msun,-0.40) kept separate below */
    double sigma_minus_mx, sigma_plus_mx;
    double vcf_mx;                      /* pow(Mx/Msun,-0.012) */
    /* --- irradiation --- */
    double Cirr, irrindex, Cirr_cold, irrindex_cold, h2rcold;
    /* --- flux --- */
    double colourfactor, emin, emax;
    /* --- T_GR constants (cpp/src/spectrum.cpp:44-68), functions of kerr and Mx only --- */
    double tgr_rg, tgr_x0, tgr_x1, tgr_x2, tgr_x3;
    /* --- wind --- */
    double windparams[4];
    /* Mdot_in the constructor of a dynamic wind sees: the BH formula on the initial F before the NS constructor shifts it
     * (cpp/src/freddi_state.cpp:74-81,156-158; T12 of SURVEY.md 9).  Read-out of windB/windC/Mdot_wind at t0 only. */
    double wind_ctor_Mdot;
    /* --- neutron star (cpp/include/ns/ns_evolution.hpp:42-67) --- */
    double R_x, redshift, R_m_min, mu_magn, R_cor, R_dead, inverse_beta, epsilon_Alfven, hot_spot_area, freqx, risco;
    double kappat[4]; /* value | in,out | in,out,par1,par2 */
    double fpparams[2]; /* romanova: par1,par2 ; geometrical: chi [rad] */
    /* --- companion star, --starflux (cpp/src/freddi_state.cpp:17-21,78-79,166-168,348-361,656-690) --- */
    double star_semiaxis;    /* the irradiation sources sit at (-semiaxis, 0, 0) */
    double star_Topt, star_albedo;
    double star_sini, star_cosi; /* of the inclination: UnitVec3(inclination, phase), cpp/src/geometry.cpp:117-124 */
    double star_t0, star_period; /* ephemeris T0 [s], orbital period [s]: phase = 2 pi (t - t0) / period */
    long long star_off;      /* offset [doubles] of this model's triangle table in the ensemble's star array */
    int32_t star_ntri, _pad2;
} fb200_model_dev_t;

#endif
