/*
 * freddi_b200 — C ABI of the B200 ensemble engine for Freddi's disc-evolution hot path.
 *
 * The reference (hombit/freddi) has no FFI for this path; its seams are C++ classes.  Each entry
 * point below names the reference interface it stands in for (paths relative to the reference
 * repository root).  Plain pointers and sizes only; no C++ or torch types cross this boundary.
 *
 *   fb200_args_t                <- FreddiArguments / FreddiNeutronStarArguments
 *                                  (cpp/include/arguments.hpp:336-354, cpp/include/ns/ns_arguments.hpp:143-165)
 *                                  with the keyword names and defaults of the Python constructor
 *                                  (cpp/pywrap/pywrap_freddi_evolution.cpp:15-83,194-224), all values CGS.
 *   fb200_ensemble_create       <- FreddiEvolution(const FreddiArguments&)                (cpp/include/freddi_evolution.hpp:50)
 *                                  FreddiNeutronStarEvolution(const FreddiNeutronStarArguments&) (cpp/include/ns/ns_evolution.hpp:228)
 *   fb200_ensemble_evolve       <- FreddiEvolution::step() repeated, i.e. the driver loops
 *                                  cpp/include/application.hpp:25-43 and EvolutionIterator (cpp/include/freddi_evolution.hpp:34-39)
 *   fb200_ensemble_rows         <- the light-curve row of FreddiFileOutput::initializeShortFields
 *                                  (cpp/src/output.cpp:197-292, cpp/src/ns/ns_output.cpp:6-19)
 *   fb200_ensemble_field        <- the array getters of FreddiState (cpp/include/freddi_state.hpp:393-400,
 *                                  cpp/pywrap/pywrap_freddi_state.cpp:35-47), NaN outside [first,last]
 *                                  like EvolutionResult (python/freddi/evolution_result.py:55-71)
 *   fb200_ensemble_state        <- FreddiState::t/i_t/first/last (cpp/include/freddi_state.hpp:309-312)
 *   fb200_ensemble_flux         <- FreddiState::flux_region<Region>(lambda) (cpp/include/freddi_state.hpp:405-407)
 *
 * Errors: the reference throws std::invalid_argument / std::runtime_error / std::logic_error at
 * construction and RadiusCollapseException during step(); nothing is thrown across this ABI.
 * Every call returns FB200_OK or a negative code, fb200_last_error() holds the message, and a
 * collapsed model is reported per model (status, rows_valid), never as a call failure.
 *
 * There is no CPU implementation behind these entry points: without a CUDA device every compute
 * call fails with FB200_ERR_NO_DEVICE.
 */
#ifndef FREDDI_B200_H
#define FREDDI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FB200_ABI_VERSION 4

/* ---- return codes ---- */
enum {
    FB200_OK = 0,
    FB200_ERR_INVALID_ARGUMENT = -1, /* reference: std::invalid_argument (bad enum value)           */
    FB200_ERR_RUNTIME = -2,          /* reference: std::runtime_error (bad parameter combination)   */
    FB200_ERR_LOGIC = -3,            /* reference: std::logic_error                                 */
    FB200_ERR_NO_DEVICE = -4,        /* no CUDA device / driver: there is no fallback               */
    FB200_ERR_CUDA = -5,             /* a CUDA runtime call failed                                  */
    FB200_ERR_UNSUPPORTED = -6,      /* outside the supported sizes (e.g. Nx too large for smem)    */
    FB200_ERR_BAD_HANDLE = -7
};

/* ---- per-model status ---- */
enum {
    FB200_MODEL_OK = 0,
    FB200_MODEL_COLLAPSED = 1, /* RadiusCollapseException "Rout <= Rin" (cpp/include/exceptions.hpp:6-11) */
    FB200_MODEL_FAILED = 2     /* Picard loop exceeded the iteration guard / non-finite state          */
};

/* ---- enumerations for the string options of the reference ---- */
enum { FB200_OPACITY_KRAMERS = 0, FB200_OPACITY_OPAL = 1 };              /* cpp/src/opacity_related.cpp:17-23 */
enum { FB200_BOUNDCOND_TEFF = 0, FB200_BOUNDCOND_TIRR = 1 };             /* cpp/src/freddi_evolution.cpp:54-68 */
enum { FB200_GRID_LOG = 0, FB200_GRID_LINEAR = 1 };                      /* cpp/src/freddi_state.cpp:36-42 */
enum { FB200_ANGDIST_PLANE = 0, FB200_ANGDIST_ISOTROPIC = 1 };           /* cpp/src/freddi_state.cpp:125-133 */
enum {                                                                     /* cpp/src/arguments.cpp:70-268, ns_arguments.cpp:148 */
    FB200_IC_POWERF = 0, FB200_IC_LINEARF = 1, FB200_IC_POWERSIGMA = 2, FB200_IC_SINEF = 3,
    FB200_IC_GAUSSF = 4, FB200_IC_QUASISTAT = 5, FB200_IC_QUASISTAT_NS = 6
};
enum {                                                                     /* cpp/src/freddi_state.cpp:94-122 */
    FB200_WIND_NO = 0,
    FB200_WIND_SS73C = 1,            /* windparams: -                       */
    FB200_WIND_CAMBIER2013 = 2,      /* windparams: kC, RIC                 */
    FB200_WIND_TESTA = 3,            /* windparams: kA                      */
    FB200_WIND_TESTB = 4,            /* windparams: kB                      */
    FB200_WIND_TESTC = 5,            /* windparams: kC                      */
    FB200_WIND_SHIELDSOSCIL1986 = 6, /* windparams: C_w, R_w                */
    FB200_WIND_JANIUK2015 = 7,       /* windparams: A_0, B_1                */
    FB200_WIND_SHIELDS1986 = 8,      /* windparams: Xi_max, T_ic, Pow       */
    FB200_WIND_WOODS1996AGN = 9,     /* windparams: C_0, T_ic               */
    FB200_WIND_WOODS1996 = 10,       /* windparams: Xi_max, T_ic, Pow       */
    FB200_WIND_TOY = 11              /* windparams: C_w                     */
};
enum { FB200_NSPROP_DUMMY = 0, FB200_NSPROP_SIBSUN2000 = 1, FB200_NSPROP_NEWT = 2 }; /* cpp/src/ns/ns_evolution.cpp:371-383 */
enum {                                                                     /* cpp/src/ns/ns_evolution.cpp:342-369 */
    FB200_FP_NO_OUTFLOW = 0, FB200_FP_PROPELLER = 1, FB200_FP_COROTATION_BLOCK = 2,
    FB200_FP_EKSI_KUTLU2010 = 3,
    FB200_FP_ROMANOVA2018 = 4, /* fpparams: par1, par2 */
    FB200_FP_GEOMETRICAL = 5   /* fpparams: chi [deg]  */
};
enum {                                                                     /* cpp/src/ns/ns_evolution.cpp:79-91 */
    FB200_KAPPAT_CONST = 0,       /* kappatparams: value               */
    FB200_KAPPAT_CORSTEP = 1,     /* kappatparams: in, out             */
    FB200_KAPPAT_ROMANOVA2018 = 2 /* kappatparams: in, out, par1, par2 */
};
enum { FB200_REDSHIFT_OFF = 0, FB200_REDSHIFT_ON = 1 };                 /* cpp/src/ns/ns_evolution.cpp:68-76 */

/*
 * One model's arguments, CGS units, keyword names of the reference's Python constructor.
 * Optional values of the reference (std::optional / Python None) are "unset" when NaN.
 * Fill with fb200_args_default() first, then override.
 */
typedef struct fb200_args {
    int32_t is_ns; /* 0: Freddi (black hole), 1: FreddiNeutronStar */
    int32_t _pad0;
    /* BasicDiskBinaryArguments, cpp/include/arguments.hpp:46-97 */
    double alpha, alphacold /*opt*/, Mx, kerr, period, Mopt, rochelobefill, Topt;
    double rin /*opt*/, rout /*opt*/, risco /*opt*/;
    /* DiskStructureArguments, cpp/include/arguments.hpp:177-239 */
    int32_t opacity, boundcond, initialcond, windtype;
    double Mdotout, Thot, Qirr2Qvishot;
    double F0 /*opt*/, Mdisk0 /*opt*/, Mdot0 /*opt*/, powerorder /*opt*/, gaussmu /*opt*/, gausssigma /*opt*/;
    double windparams[4]; /* meaning per windtype, see enum */
    /* SelfIrradiationArguments, cpp/include/arguments.hpp:242-265 */
    double Cirr, irrindex, Cirrcold, irrindexcold, h2rcold;
    int32_t angulardistdisk, angulardistns;
    /* FluxArguments, cpp/include/arguments.hpp:268-304 (lambdas/passbands are ensemble-wide, see fb200_config_t) */
    double colourfactor, emin, emax, staralbedo, inclination, ephemerist0, distance;
    /* CalculationArguments, cpp/include/arguments.hpp:307-333 */
    double inittime, time, tau /*opt*/, eps;
    uint32_t Nx;
    int32_t gridscale;
    int32_t starlod; /* level of detail of the companion star (20 * 4^starlod triangles), used with fb200_config_t::star_flux */
    int32_t nsprop;
    /* NeutronStarArguments, cpp/include/ns/ns_arguments.hpp:16-67 (ignored when !is_ns) */
    double freqx /*opt*/, Rx /*opt*/, Bx, hotspotarea, epsilonAlfven, inversebeta, Rdead;
    int32_t fptype, kappattype, nsgravredshift, _pad1;
    double fpparams[2];
    double kappatparams[4];
} fb200_args_t;

/* Ensemble-wide configuration: the outputs every model of the ensemble produces. */
#define FB200_MAX_LAMBDAS 16
#define FB200_MAX_PASSBANDS 4
#define FB200_MAX_PASSBAND_SAMPLES 2048 /* tabulated points per passband */
typedef struct fb200_config {
    int32_t device;         /* CUDA device ordinal                                                        */
    int32_t n_lambdas;      /* Fnu columns, FluxArguments::lambdas (cpp/include/arguments.hpp:286)        */
    double lambdas[FB200_MAX_LAMBDAS]; /* cm                                                               */
    int32_t cold_disk;      /* add Fnu{k}_cold columns (--colddiskflux, cpp/src/output.cpp:226-233)       */
    int32_t record_F;       /* keep F(h) of every row in HBM so fb200_ensemble_field can serve any row    */
    int32_t compute_lx;     /* 1 (reference behaviour): Lx/Fx by the adaptive quadrature; 0: columns NaN  */
    int32_t host_threads;   /* threads for host-side argument resolution, 0 = all cores                   */
    int32_t threads_per_model; /* reserved (the CTA size is a build-time constant)                       */
    int32_t exact_lx_walk;  /* 0 (default): rings whose X-ray band emission is provably negligible against the
                             * hottest ring's are not integrated (changes Lx by < 1e-16 relative, bound per row),
                             * and a ring's band integral is read from the table of the adaptive walk's own result
                             * (csrc/lx_table.h: the walk is a function of h emin / kT alone; <= 2e-12 per ring);
                             * 1: walk every ring like the reference does (Spectrum::Planck_nu1_nu2)            */
    /* Tabulated passbands, FluxArguments::passbands (cpp/include/arguments.hpp:287): one Fnu<name> column each
     * (+ Fnu<name>_cold with cold_disk) after the Fnu{k} columns, cpp/src/output.cpp:258-271.  The arrays are what
     * EnergyPassband holds (cpp/include/passband.hpp:33-36): wavelengths [cm] and transmission factors, for the
     * photon-counter detector of the CLI already multiplied by the wavelength (cpp/src/passband.cpp:35-37).  They are
     * read during fb200_ensemble_create only. */
    int32_t n_passbands;
    int32_t passband_len[FB200_MAX_PASSBANDS];
    const double* passband_lambdas[FB200_MAX_PASSBANDS];
    const double* passband_transmissions[FB200_MAX_PASSBANDS];
    /* --starflux (cpp/src/options.cpp:333): after every Fnu column (and its _cold one) add Fnu*_star, Fnu*_star_min,
     * Fnu*_star_max — the irradiated companion star at the current orbital phase and at the two conjunctions
     * (cpp/src/output.cpp:234-255,272-291).  The Roche-lobe surface is triangulated on the host at create
     * (20 * 4^starlod triangles per distinct Mopt/Mx, rochelobefill, semi-axis).
     * 2: build the geometry only, for fb200_ensemble_flux_rows(region 2) — no columns, no per-row work. */
    int32_t star_flux;
    /* A model's time steps are handed out in chunks to whichever CTA is free; a chunk waits for the model's previous chunk.
     * If that one is not published within wait_timeout_ms (its CTA can only have died) the model is reported
     * FB200_MODEL_FAILED with the rows finished so far, instead of hanging the launch.  0: 10 s. */
    int32_t wait_timeout_ms;
} fb200_config_t;

/* Row layout: cpp/src/output.cpp:197-214, then per wavelength Fnu{k}[, Fnu{k}_cold][, Fnu{k}_star, _star_min, _star_max],
 * then the same group per passband (cpp/src/output.cpp:219-291), then cpp/src/ns/ns_output.cpp:6-19 for NS. */
enum {
    FB200_COL_T = 0,        /* days */
    FB200_COL_MDOT = 1, FB200_COL_MDISK = 2,
    FB200_COL_RHOT = 3,     /* Rsun */
    FB200_COL_SIGMAOUT = 4, FB200_COL_KIRROUT = 5, FB200_COL_H2R = 6, FB200_COL_TEFFOUT = 7,
    FB200_COL_TIRROUT = 8, FB200_COL_QIRR2QVISOUT = 9,
    FB200_COL_TPHXMAX = 10, /* keV */
    FB200_COL_LX = 11, FB200_COL_LBOL = 12, FB200_COL_FX = 13, FB200_COL_FBOL = 14,
    FB200_N_BH_COLS = 15,
    /* NS extras follow the Fnu columns, in this order */
    FB200_NSCOL_RIN = 0, FB200_NSCOL_ETANS = 1, FB200_NSCOL_LXNS = 2, FB200_NSCOL_LBOLNS = 3,
    FB200_NSCOL_FXNS = 4, FB200_NSCOL_FBOLNS = 5, FB200_NSCOL_THOTSPOT = 6 /* keV */,
    FB200_NSCOL_FPIN = 7, FB200_NSCOL_FMAGNIN = 8, FB200_NSCOL_FIN = 9,
    FB200_N_NS_COLS = 10
};

/* Radial fields served by fb200_ensemble_field (cpp/pywrap/pywrap_freddi_state.cpp:35-47, ns: pywrap_freddi_evolution.cpp:317-319) */
enum {
    FB200_FIELD_H = 0, FB200_FIELD_R = 1, FB200_FIELD_F = 2, FB200_FIELD_W = 3, FB200_FIELD_SIGMA = 4,
    FB200_FIELD_TPH = 5, FB200_FIELD_TPH_VIS = 6, FB200_FIELD_TIRR = 7, FB200_FIELD_KIRR = 8,
    FB200_FIELD_HEIGHT = 9, FB200_FIELD_QX = 10, FB200_FIELD_TPH_X = 11,
    FB200_FIELD_WINDA = 12, FB200_FIELD_WINDB = 13, FB200_FIELD_WINDC = 14,
    FB200_FIELD_FMAGN = 15, FB200_FIELD_DFMAGN_DH = 16, FB200_FIELD_D2FMAGN_DH2 = 17,
    FB200_N_FIELDS = 18
};

/* Per-model scalar constants resolved at construction (FreddiState getters, cpp/include/freddi_state.hpp:286-295) */
typedef struct fb200_model_info {
    double GM, eta, R_g, cosi, distance, cosiOverD2, tau, h_in, h_out, rin, rout, risco;
    double F0, Mdisk0, Mdot0;                 /* as normalised by the initial condition (cpp/src/arguments.cpp:57-269) */
    double mu_magn, R_cor, R_dead, inverse_beta, R_x, redshift; /* NS, cpp/include/ns/ns_evolution.hpp:185-198 */
    int64_t Nx, Nt, first0;
} fb200_model_info_t;

typedef struct fb200_ensemble fb200_ensemble_t;

/* --- library --- */
int fb200_abi_version(void);
const char* fb200_last_error(void); /* thread-local message of the last failing call */
int fb200_device_count(void);       /* number of CUDA devices visible, <0 on driver error */
/* Measured FP64 peak of `device` [TFLOP/s, 2 flop per DFMA]: a register-resident DFMA micro-kernel timed with
 * CUDA events.  It is the denominator of the roofline this path is reported against (the path is bound by
 * the FP64 pipe, not by HBM). */
int fb200_measure_fp64_peak(int device, double* tflops);

/* Self-test of the band-integral table (freddi_b200/csrc/lx_table.h): Lx sums, per ring, the integral of B_nu over [emin, emax]
 * that the reference takes with an adaptive Cash-Karp walk (Spectrum::Planck_nu1_nu2, cpp/src/spectrum.cpp:26-37).  For a fixed
 * band ratio that walk's result is a piecewise-analytic function of x1 = h nu1 / kT; the library tabulates it on first use and the
 * kernels look it up.  This entry builds (or fetches) the table of the band [nu1, nu2] (Hz) and compares it on the host with the
 * direct walk at `n` temperatures, log-uniform in x1 over [x_lo, x_hi]: the largest relative deviation, the number of points
 * whose try_step count differs, the x1 of the largest deviation, the table's number of sub-pieces.  FB200_ERR_UNSUPPORTED when no
 * table can be built for this ratio (the kernels then walk).  Needs no device. */
int fb200_lx_table_selftest(double nu1, double nu2, double x_lo, double x_hi, long n, unsigned long long seed, double* max_rel_err,
                            long* trials_mismatches, double* worst_x1, int* n_sub);

/* Self-test: trapz(x, y, first, last) of the reference (cpp/src/util.cpp:9-19) evaluated by the CTA-wide weighted sum the
 * kernels use for every radial integral (Mdisk, Lx, Fnu, Mdot_wind), on caller data: 2 <= n <= 1024 host doubles.  The
 * reference's known answer (cpp/test/util.cpp:31-41) is asserted through this entry. */
int fb200_selftest_trapz(int device, const double* x, const double* y, size_t n, size_t first, size_t last, double* out);

/* --- arguments --- */
void fb200_args_default(fb200_args_t* args, int is_ns); /* defaults of pywrap_freddi_evolution.cpp:33-83,203-224 */
void fb200_config_default(fb200_config_t* cfg);

/* Host-only resolution of one model's arguments (no device needed): the constants, the grid h[Nx] and the
 * initial F[Nx] that fb200_ensemble_create uploads.  Stands in for constructing the reference's Arguments
 * blocks + FreddiState (cpp/src/arguments.cpp:30-351, cpp/src/freddi_state.cpp:13-71).  h / F may be NULL;
 * when given they must hold args->Nx doubles.  Returns the error class the reference would throw. */
int fb200_args_resolve(const fb200_args_t* args, fb200_model_info_t* info, double* h, double* F);

/* --- ensemble life cycle --- */
/* Resolves every model's arguments on the host (ISCO, tidal radius, opacity constants, grid h, initial
 * F(h), static wind and magnetic-torque arrays: cpp/src/freddi_state.cpp:13-71, cpp/src/arguments.cpp:57-351,
 * cpp/src/ns/ns_evolution.cpp:45-139,322-340), uploads them and allocates row storage for Nt+1 rows.
 * All models of one ensemble share is_ns and Nx; time/tau may differ (rows are padded to the largest Nt+1).
 * On a bad argument returns the reference's error class and *bad_model = index of the offending model. */
int fb200_ensemble_create(const fb200_args_t* args, size_t n_models, const fb200_config_t* cfg,
                          fb200_ensemble_t** out, size_t* bad_model);
void fb200_ensemble_destroy(fb200_ensemble_t* e);

size_t fb200_ensemble_n_models(const fb200_ensemble_t* e);
size_t fb200_ensemble_n_cols(const fb200_ensemble_t* e);  /* columns per row                     */
size_t fb200_ensemble_n_rows(const fb200_ensemble_t* e);  /* rows per model = max(Nt)+1          */
size_t fb200_ensemble_nx(const fb200_ensemble_t* e);
int fb200_ensemble_info(const fb200_ensemble_t* e, fb200_model_info_t* out /*[n_models]*/);

/* Put every model back to its initial state (t = inittime, F = F(t=0), first/last, no rows) from the copy of the resolved
 * initial conditions the ensemble keeps in HBM: the equivalent of constructing the same FreddiEvolution objects again
 * (cpp/include/freddi_evolution.hpp:50) without repeating the argument resolution. */
int fb200_ensemble_reset(fb200_ensemble_t* e);

/* Advance every model by n_steps (clipped per model at its Nt; n_steps<0: to the end).  Row i_t is
 * written when state i_t is reached; row 0 is written by the first call.  Blocks until done. */
int fb200_ensemble_evolve(fb200_ensemble_t* e, int64_t n_steps);
/* Same, but only enqueues on the ensemble's stream; pair with fb200_ensemble_sync. */
int fb200_ensemble_evolve_async(fb200_ensemble_t* e, int64_t n_steps);
int fb200_ensemble_sync(fb200_ensemble_t* e);
/* Device time of the last evolve call in milliseconds (CUDA events on the ensemble's stream). */
int fb200_ensemble_last_evolve_ms(fb200_ensemble_t* e, float* ms);
/* Number of kernel launches issued by the last evolve call. */
int fb200_ensemble_last_launches(const fb200_ensemble_t* e);

/* Copy rows to host memory: out[model][row][col], n_rows x n_cols per model; rows >= rows_valid are NaN. */
int fb200_ensemble_rows(fb200_ensemble_t* e, double* out, size_t out_len);
/* Device pointer of the same buffer (for callers that keep results in HBM). */
int fb200_ensemble_rows_device(fb200_ensemble_t* e, const double** dptr);
/* Per-model status, number of valid rows, Picard iterations summed over steps. */
int fb200_ensemble_status(fb200_ensemble_t* e, int32_t* status, int32_t* rows_valid, int64_t* picard_iters);
/* Work counters per model, summed over the steps done so far: Picard iterations of the solver and try_step
 * calls of the Lx quadrature (6 Planck evaluations each in the reference).  For the roofline arithmetic. */
int fb200_ensemble_counters(fb200_ensemble_t* e, int64_t* picard_iters, int64_t* lx_trial_steps);
/* All work counters per model: out[model][FB200_N_WORK_COUNTERS].  RING_STEPS = sum over the implicit solves of the
 * unknowns (last - first) they had, RING_ITERS = the same weighted by the Picard iterations of each solve, POW_EVALS =
 * closure pow() calls the Picard loops made (the incremental K update replaces most): what the kernel really did, for
 * the roofline arithmetic next to the reference's operation count. */
enum { FB200_WORK_PICARD_ITERS = 0, FB200_WORK_LX_TRIALS = 1, FB200_WORK_RING_STEPS = 2, FB200_WORK_RING_ITERS = 3,
       FB200_WORK_POW_EVALS = 4, FB200_WORK_ROW_RINGS = 5 /* sum over the rows of last - first + 1 */, FB200_N_WORK_COUNTERS = 6 };
int fb200_ensemble_work_counters(fb200_ensemble_t* e, int64_t* out, size_t out_len);
/* Current t [s], i_t, first, last per model (any pointer may be NULL). */
int fb200_ensemble_state(fb200_ensemble_t* e, double* t, int64_t* i_t, int64_t* first, int64_t* last);
/* State of a recorded row (requires record_F): first/last per model at that row. */
int fb200_ensemble_row_bounds(fb200_ensemble_t* e, int64_t row, int64_t* first, int64_t* last);

/* Radial field of every model: out[model][Nx], exactly as the reference getter returns it (zeros below
 * `first`, cold-zone values beyond `last`; the NaN padding of EvolutionResult is done by the Python layer).
 * row<0: the current state; row>=0: that recorded row (requires cfg.record_F). */
int fb200_ensemble_field(fb200_ensemble_t* e, int field, int64_t row, double* out, size_t out_len);
/* Spectral flux density of the hot (region 0) or cold (region 1) disc at wavelengths lambdas[cm]:
 * out[model][n_lambdas], evaluated on the state of `row` (row<0: current). */
int fb200_ensemble_flux(fb200_ensemble_t* e, int region, int64_t row, const double* lambdas, size_t n_lambdas,
                        double* out, size_t out_len);
/* The same for a range of recorded rows in ONE launch (work items = (model, row) pairs): out[model][n_rows][Nx], with
 * pad_nan != 0 NaN outside [first, last] — the (Nt+1, Nx) arrays of EvolutionResult (python/freddi/evolution_result.py:55-71)
 * without a call per state.  Requires cfg.record_F. */
int fb200_ensemble_field_rows(fb200_ensemble_t* e, int field, int64_t row_begin, int64_t n_rows, int pad_nan, double* out, size_t out_len);
/* Spectral flux densities of a range of rows in one launch: out[model][n_rows][n_lambdas].  region 0 hot, 1 cold,
 * 2 the irradiated companion star seen at orbital phase phases[row - row_begin] [rad] (FreddiState::flux_star(lambda, phase),
 * cpp/src/freddi_state.cpp:355-357; needs cfg.star_flux).  row_begin < 0 with n_rows = 1: the current state. */
int fb200_ensemble_flux_rows(fb200_ensemble_t* e, int region, int64_t row_begin, int64_t n_rows, const double* lambdas, size_t n_lambdas,
                             const double* phases, double* out, size_t out_len);
/* Mdot_wind of FreddiState::Mdot_wind (cpp/src/freddi_state.cpp:364-383), per model, on the state of `row`. */
int fb200_ensemble_mdot_wind(fb200_ensemble_t* e, int64_t row, double* out, size_t out_len);

/* --- device-resident results (zero copy) --- */
/* Pointers into the ensemble's HBM for consumers that stay on the GPU (the reference's Python layer hands out host copies,
 * cpp/pywrap/converters.cpp:39-48; python/freddi/evolution_result.py:55-71 stacks them): ROWS [n_models][n_rows][n_cols],
 * F [n_models][Nx] (current state), H [n_models][Nx], and with record_F FSNAP [n_models][n_rows][Nx] (NaN where never written)
 * and BOUNDS int32 [n_models][n_rows][2] (first, last per row).  Valid until the ensemble is destroyed; synchronise with
 * fb200_ensemble_sync before reading after an asynchronous evolve.  *device = the CUDA ordinal the memory lives on. */
enum { FB200_DEVPTR_ROWS = 0, FB200_DEVPTR_F = 1, FB200_DEVPTR_H = 2, FB200_DEVPTR_FSNAP = 3, FB200_DEVPTR_BOUNDS = 4 };
int fb200_ensemble_device_ptr(fb200_ensemble_t* e, int which, const void** dptr, int32_t* device);

/* --- state in / out: checkpoint and resume --- */
/* The copyable state of the reference (FreddiState's copy constructor, cpp/src/freddi_state.cpp:84-91: the structure is shared,
 * CurrentState is copied).  A checkpoint is an opaque blob holding every model's CurrentState (t, i_t, first, last, F_in,
 * Mdot_in_prev, status, work counters, the last wind update), F, and the closure K carried with it, optionally the rows written
 * so far; loaded into an ensemble created from the same arguments it makes the continuation bit-identical to an uninterrupted
 * run. */
size_t fb200_ensemble_checkpoint_bytes(const fb200_ensemble_t* e, int with_rows);
int fb200_ensemble_checkpoint_save(fb200_ensemble_t* e, void* buf, size_t len, int with_rows);
int fb200_ensemble_checkpoint_load(fb200_ensemble_t* e, const void* buf, size_t len);
/* Explicit state: F[n_models][Nx] and the scalars of CurrentState per model (cpp/include/freddi_state.hpp:242-258); any pointer
 * may be NULL (get: not wanted; set: keep).  set_state makes every model alive again at row i_t (row 0 is written from the
 * given state when i_t == 0) and voids the carried closure, so a run continued from an explicit state agrees with an
 * uninterrupted one to rounding (~1e-16 per step), not bit for bit: use a checkpoint for that. */
int fb200_ensemble_get_state(fb200_ensemble_t* e, double* F, double* t, int64_t* i_t, int64_t* first, int64_t* last, double* F_in,
                             double* Mdot_in_prev);
int fb200_ensemble_set_state(fb200_ensemble_t* e, const double* F, const double* t, const int64_t* i_t, const int64_t* first,
                             const int64_t* last, const double* F_in, const double* Mdot_in_prev);

/* ------------------------------------------------------------------------------------------------------------------
 * In-process multi-GPU: one handle over several devices.  Stands in for running the reference's driver (cpp/main.cpp:7-12
 * -> cpp/include/application.hpp:17-43: construct, loop over step(), dump a row per step) once per model of a sweep.
 * Model k of the caller lives on shard k mod G (G = number of devices given; a device may be listed more than once) at
 * local index k / G; shards never exchange data.  devices == NULL or n_devices <= 0: all visible devices.  Every output
 * is in the caller's model order.  No torch, no MPI: plain host threads and CUDA streams.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct fb200_sharded fb200_sharded_t;
int fb200_sharded_create(const fb200_args_t* args, size_t n_models, const fb200_config_t* cfg, const int32_t* devices, int32_t n_devices,
                         fb200_sharded_t** out, size_t* bad_model); /* cfg->device is ignored */
void fb200_sharded_destroy(fb200_sharded_t* h);
int32_t fb200_sharded_n_shards(const fb200_sharded_t* h);
fb200_ensemble_t* fb200_sharded_shard(fb200_sharded_t* h, int32_t k); /* the single-device handle of shard k (owned by h) */
size_t fb200_sharded_n_models(const fb200_sharded_t* h);
size_t fb200_sharded_n_rows(const fb200_sharded_t* h);
size_t fb200_sharded_n_cols(const fb200_sharded_t* h);
size_t fb200_sharded_nx(const fb200_sharded_t* h);
int fb200_sharded_info(fb200_sharded_t* h, fb200_model_info_t* out /*[n_models]*/);
int fb200_sharded_reset(fb200_sharded_t* h);
int fb200_sharded_evolve(fb200_sharded_t* h, int64_t n_steps);       /* all devices at once; blocks until all are done */
int fb200_sharded_evolve_async(fb200_sharded_t* h, int64_t n_steps);
int fb200_sharded_sync(fb200_sharded_t* h);
int fb200_sharded_last_evolve_ms(fb200_sharded_t* h, float* ms);     /* the longest launch over the devices */
/* out[n_models][n_rows][n_cols]: every device copies its models into its strided slice of the one buffer, all at once.
 * A buffer from fb200_host_alloc (pinned) is copied to at full PCIe rate; pageable memory works, slower. */
int fb200_sharded_rows(fb200_sharded_t* h, double* out, size_t out_len);
int fb200_sharded_status(fb200_sharded_t* h, int32_t* status, int32_t* rows_valid, int64_t* picard_iters);
int fb200_sharded_state(fb200_sharded_t* h, double* t, int64_t* i_t, int64_t* first, int64_t* last);
int fb200_sharded_work_counters(fb200_sharded_t* h, int64_t* out, size_t out_len);
int fb200_sharded_field(fb200_sharded_t* h, int field, int64_t row, double* out, size_t out_len);
int fb200_sharded_flux(fb200_sharded_t* h, int region, int64_t row, const double* lambdas, size_t n_lambdas, double* out, size_t out_len);

/* One-shot sweep — the whole of application.hpp:17-43 for every model in one call: argument structs in, light curves out.
 * Construction is streamed: the evolve kernel of every device is launched first, host threads resolve the models in upload
 * order into pinned staging, each uploaded group of models is picked up by the running kernel as its watermark arrives, and
 * the rows of finished groups are copied to `rows` while later groups still run — argument resolution, allocation, H2D and
 * most of the D2H hide behind the device time.  rows[n_models][n_rows][n_cols] with n_rows, n_cols from fb200_sweep_plan;
 * status / rows_valid / picard_iters / t_end (FreddiState::t() where the model stopped [s]) [n_models] may be NULL.
 * cfg->record_F is ignored (rows only). */
typedef struct fb200_sweep_stats {
    double total_ms;        /* the whole call                                                                   */
    double device_ms_max;   /* longest evolve launch over the devices (CUDA events; includes its waits for uploads) */
    double first_launch_ms; /* entry -> every device's kernel launched                                          */
    double resolve_ms;      /* entry -> last model resolved on the host                                         */
    double uploads_ms;      /* entry -> last upload complete                                                    */
    int64_t model_steps;    /* sum over the models of the steps they took (rows_valid - 1)                      */
    int32_t n_devices, launches, n_rows, n_cols;
} fb200_sweep_stats_t;
int fb200_sweep_plan(const fb200_args_t* args, size_t n_models, const fb200_config_t* cfg, size_t* n_rows, size_t* n_cols);
int fb200_sweep_run(const fb200_args_t* args, size_t n_models, const fb200_config_t* cfg, const int32_t* devices, int32_t n_devices,
                    double* rows, size_t rows_len, int32_t* status, int32_t* rows_valid, int64_t* picard_iters, double* t_end,
                    fb200_sweep_stats_t* stats, size_t* bad_model);

/* Pinned (page-locked, portable) host memory for result buffers; fb200_trim gives back the device memory (a stream-ordered pool
 * per device) and the pinned staging blocks the library keeps between ensembles (allocation costs milliseconds, and cudaFree
 * would synchronise the device under a running sweep; FB200_NO_CACHE=1: nothing is kept). */
int fb200_host_alloc(size_t bytes, void** p);
void fb200_host_free(void* p);
void fb200_trim(void);

#ifdef __cplusplus
}
#endif
#endif /* FREDDI_B200_H */
