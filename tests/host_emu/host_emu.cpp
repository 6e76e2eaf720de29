// TEST INFRASTRUCTURE — not part of the product, never linked into libfreddi_b200.so.
//
// Runs DiscEngine (freddi_b200/csrc/disc_engine.h — the very code the CUDA kernel executes) under a
// sequential execution policy: points() becomes a loop over the radial points, barriers are no-ops and
// reductions are plain loops.  This lets the CPU test suite check the step/row logic against the oracle
// without a GPU.  It cannot see races or barrier mistakes; the `-m gpu` tests and compute-sanitizer do.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../freddi_b200/csrc/disc_engine.h"
#include "../../freddi_b200/csrc/resolve.hpp"
#include "../../freddi_b200/csrc/star.hpp"

using namespace fb200;

namespace {

struct HostEx {
    typedef PtrArrays Arrays;
    static constexpr int NT = FB200_NX_MAX;
    double scal[8] = {0};
    double sums[FB200_MAXQ] = {0};
    double lampref[FB200_MAX_LAMBDAS] = {0};
    const double *gA = nullptr, *gB = nullptr, *gCs = nullptr, *gFm = nullptr, *gdFm = nullptr;

    int cur = 0;  // point the lambda is running for: the host policy has one register slot per point
    template <class F> void points(F&& f) { for (int i = 0; i < NT; ++i) { cur = i; f(i, 0); } }
    template <class F> void points_ilp(F&& f) { points(f); }
    double reduce_max_of(double v) const { return v; }  // the sequential policy accumulated it over all points already
    void sync() {}
    void tick(int) {}
    static constexpr bool kWarpPcr = false;  // the plain PCR loop; the device's hybrid does the same arithmetic level by level
    int thread() const { return 0; }
    double shfl_up(double v, int) const { return v; }
    double shfl_down(double v, int) const { return v; }
    template <class F> void blocks(int n, F&& f) { for (int b = 0; b < n; ++b) f(b); }
    template <class F> void reduced(int n, F&& f) { for (int q = 0; q < n; ++q) f(q); }
    double& scalar(int k) { return scal[k]; }
    long long counters[FB_CNT_N] = {0};
    void count(int c, long long v) { counters[c] += v; }
    void note_pow_fallbacks(int n) { counters[FB_CNT_POW_FALLBACKS] += n; }
    void note_lx_trials(int n) { counters[FB_CNT_LX_TRIALS] += n; }
    bool any_of(bool p) const { return p; }  // the sequential policy accumulated the predicate's operand over all points already
    WindRecord wind{0., 1, 0};
    void note_wind_update(double Mdot, int first, int last) { wind = WindRecord{Mdot, first, last}; }
    double sum(int q) const { return sums[q]; }
    double lambda_pref(int k) const { return lampref[k]; }
    double windA(int i) const { return gA ? gA[i] : 0.; }
    double windB(int i) const { return gB ? gB[i] : 0.; }
    double windC_static(int i) const { return gCs ? gCs[i] : 0.; }
    bool has_windC() const { return gCs != nullptr; }
    double Fmagn(int i) const { return gFm ? gFm[i] : 0.; }
    double dFmagn_dh(int i) const { return gdFm ? gdFm[i] : 0.; }
    std::vector<double> starT;
    double starsum[3 * (FB200_MAX_LAMBDAS + FB200_MAX_PASSBANDS)] = {0};
    double& star_T(int t) { return starT[t]; }
    double& star_sum(int k) { return starsum[k]; }
    template <class F> void tri_points(int n, F&& f) { for (int t = 0; t < n; ++t) f(t); }
    template <class F> void reduce_sums3(int n, F&& f) {
        double s0 = 0., s1 = 0., s2 = 0.;
        for (int t = 0; t < n; ++t) { double a, b, c; f(t, a, b, c); s0 += a; s1 += b; s2 += c; }
        sums[0] = s0; sums[1] = s1; sums[2] = s2;
    }
    template <class F> double reduce_max(F&& f) {
        double m = -INFINITY;
        for (int i = 0; i < NT; ++i) { cur = i; const double v = f(i, 0); if (v == v) m = std::fmax(m, v); }
        return m;
    }
    template <class F> int reduce_max_int(F&& f) { int m = 0x80000000; for (int i = 0; i < NT; ++i) { cur = i; m = std::max(m, f(i, 0)); } return m; }
    template <class F> int reduce_min_int(F&& f) { int m = 0x7fffffff; for (int i = 0; i < NT; ++i) { cur = i; m = std::min(m, f(i, 0)); } return m; }
    template <class F> void reduce_sums(int nq, F&& f) {
        for (int q = 0; q < nq; ++q) { double s = 0.; for (int i = 0; i < NT; ++i) { cur = i; s += f(i, 0, q); } sums[q] = s; }
    }
};

}  // namespace

static long long g_last_lx_trials = 0;
static std::vector<PassbandTable> g_passbands;  // set by emu_set_passbands for the runs that follow
static int g_lx_table = 1;                       // emu_set_lx_table(0): every ring walks (as exact_lx_walk)
static int g_star_flux = 0;                      // set by emu_set_star_flux for the runs that follow

extern "C" {

// Same contract as ref_set_passbands (oracle/ref_driver.cpp); returns 0 or the resolver's error code.
int emu_set_passbands(int n, const int* lens, const double* lambdas, const double* transmissions, double* t_dnu) {
    g_passbands.clear();
    size_t off = 0;
    for (int p = 0; p < n; ++p) {
        PassbandTable t;
        std::string msg;
        const int rc = prepare_passband(lambdas + off, transmissions + off, lens[p], t, msg);
        if (rc != FB200_OK) { g_passbands.clear(); return rc; }
        off += lens[p];
        if (t_dnu) t_dnu[p] = t.t_dnu;
        g_passbands.push_back(t);
    }
    return 0;
}

// Same contract as ref_run (oracle/ref_driver.cpp): returns the number of valid rows, <0 on a resolve error.
int64_t emu_run(const fb200_args_t* a, const double* lambdas, int n_lambdas, int cold_disk, double* rows, int64_t rows_cap,
                double* Fsnap, int64_t* first, int64_t* last, int64_t* picard, char* err, size_t err_len) {
    Resolved res;
    std::string msg;
    const int rc = resolve(*a, res, msg);
    if (rc != FB200_OK) {
        if (err && err_len) std::snprintf(err, err_len, "%s", msg.c_str());
        return rc;
    }
    const fb200_model_dev_t& M = res.dev;
    const int Nx = M.Nx;
    OutCfg cfg{};
    cfg.n_lambdas = n_lambdas;
    cfg.cold_disk = cold_disk;
    cfg.compute_lx = 1;
    cfg.lx_prune = 1;
    if (g_lx_table) {  // as build_ensembles (construct.cu): the table of this model's band
        size_t len = 0;
        cfg.lx_table = lx_table_get(a->emax / a->emin, &len);
        res.dev.lx_tab = cfg.lx_table ? 1 : 0;
    }
    const int npb = int(g_passbands.size());
    std::vector<double> pb_a, pb_w;
    cfg.n_passbands = npb;
    for (int p = 0; p < npb; ++p) {
        cfg.pb_off[p] = int(pb_a.size());
        pb_a.insert(pb_a.end(), g_passbands[p].a.begin(), g_passbands[p].a.end());
        pb_w.insert(pb_w.end(), g_passbands[p].w.begin(), g_passbands[p].w.end());
        cfg.pb_tdnu[p] = g_passbands[p].t_dnu;
    }
    for (int p = npb; p < FB200_MAX_PASSBANDS + 2; ++p) cfg.pb_off[p] = int(pb_a.size());
    cfg.pb_a = pb_a.data();
    cfg.pb_w = pb_w.data();
    cfg.star_flux = g_star_flux;
    StarGeometry geom;
    std::vector<double> star_packed;
    if (g_star_flux) {
        std::string serr;
        const int src = build_star_geometry(res.dev.star_semiaxis, a->Mopt / a->Mx, a->rochelobefill, a->starlod, geom, serr);
        if (src != FB200_OK) { if (err && err_len) std::snprintf(err, err_len, "%s", serr.c_str()); return src; }
        star_packed = geom.packed();
        res.dev.star_off = 0;
        res.dev.star_ntri = geom.n_tri;
        cfg.star_data = star_packed.data();
    }
    cfg.n_cols = FB200_N_BH_COLS + (n_lambdas + npb) * ((cold_disk ? 2 : 1) + (g_star_flux ? 3 : 0)) + (M.is_ns ? FB200_N_NS_COLS : 0);
    cfg.n_rows = M.Nt + 1;
    for (int k = 0; k < n_lambdas; ++k) cfg.lambdas[k] = lambdas[k];

    const int N = FB200_NX_MAX, NP = FB200_NPAD, NR = FB200_NRED_MAX;
    std::vector<double> pool(18 * size_t(N) + 7 * size_t(NP) + 6 * size_t(NR), 0.);
    double* p = pool.data();
    auto take = [&](size_t n) { double* q = p; p += n; return q; };
    double* hbuf = take(N);
    PtrArrays S;
    S.h = hbuf; S.R = take(N); S.F = take(N); S.hn = take(N);
    S.ca = take(N); S.cb = take(N); S.cc = take(N); S.frac = take(N);
    S.c1 = take(N); S.hm175 = take(N); S.hotR = take(N); S.Gb = take(N);
    S.sgA = take(N); S.qB = take(N); S.tv4 = take(N);
    S.s_cc = take(N); S.s_frac = take(N); S.s_K = take(N);
    S.s_l = take(NP); S.s_u = take(NP); S.s_d = take(NP); S.s_rhs = take(NP); S.p_a = take(NP); S.p_c = take(NP); S.p_g = take(NP);
    S.lu[0] = reinterpret_cast<double2*>(take(2 * NR)); S.lu[1] = reinterpret_cast<double2*>(take(2 * NR));
    S.rr[0] = take(NR); S.rr[1] = take(NR);
    S.Tph = S.s_l; S.Tirr = S.s_u; S.Sigma = S.s_d; S.Tvis = S.s_rhs; S.TphX = S.p_a; S.aux = S.p_c; S.wq = S.p_g;
    for (int i = 0; i < Nx; ++i) { hbuf[i] = res.h[i]; S.F[i] = res.F[i]; }

    HostEx ex;
    ex.starT.assign(size_t(geom.n_tri) + 1, 0.);
    for (int k = 0; k < n_lambdas; ++k) ex.lampref[k] = FB_DOUBLE_H_C2 / p5(lambdas[k]);
    std::vector<double> wCs;
    if (!res.wC.empty() || (M.is_ns && M.inverse_beta != 0.)) {
        wCs = res.wC.empty() ? std::vector<double>(Nx, 0.) : res.wC;
        if (!res.d2Fmagn_dh2.empty()) for (int i = 0; i < Nx; ++i) wCs[i] += res.d2Fmagn_dh2[i];
        ex.gCs = wCs.data();
    }
    if (!res.wA.empty()) ex.gA = res.wA.data();
    if (!res.wB.empty()) ex.gB = res.wB.data();
    if (!res.Fmagn.empty()) ex.gFm = res.Fmagn.data();
    if (!res.dFmagn_dh.empty()) ex.gdFm = res.dFmagn_dh.data();

    ModelState st{};
    st.t = M.t0; st.F_in = M.F_in0; st.Mdot_in_prev = -INFINITY; st.first = M.first0; st.last = Nx - 1;
    DiscEngine<HostEx> eng(ex, M, cfg, S, st);
    eng.init_points();

    const int nc = cfg.n_cols;
    std::vector<double> rowbuf(nc);
    std::vector<double> Fbuf(N);
    int bounds[2];
    int64_t r = 0;
    for (; r <= M.Nt && r < rows_cap; ++r) {
        bool ok;
        const long long before = ex.counters[FB_CNT_PICARD];
        if (r == 0) ok = eng.structure_and_row(false, rowbuf.data(), Fbuf.data(), bounds);
        else ok = eng.step_and_row(rowbuf.data(), Fbuf.data(), bounds);
        if (!ok) break;
        std::memcpy(rows + r * nc, rowbuf.data(), sizeof(double) * nc);
        if (Fsnap) std::memcpy(Fsnap + r * Nx, Fbuf.data(), sizeof(double) * Nx);
        if (first) first[r] = bounds[0];
        if (last) last[r] = bounds[1];
        if (picard) picard[r] = ex.counters[FB_CNT_PICARD] - before;
    }
    g_last_lx_trials = ex.counters[FB_CNT_LX_TRIALS];
    return r;
}

// Companion-star triangles as the product's host code builds them (star.cpp): [n][7] centre, normal, area; returns n or <0.
long emu_star_triangles(const fb200_args_t* a, double* out, long cap) {
    Resolved res;
    std::string msg;
    if (resolve(*a, res, msg) != FB200_OK) return -1;
    StarGeometry g;
    if (build_star_geometry(res.dev.star_semiaxis, a->Mopt / a->Mx, a->rochelobefill, a->starlod, g, msg) != FB200_OK) return -2;
    for (long i = 0; i < g.n_tri && i < cap; ++i) {
        double* o = out + 7 * i;
        o[0] = g.cx[i]; o[1] = g.cy[i]; o[2] = g.cz[i]; o[3] = g.nx[i]; o[4] = g.ny[i]; o[5] = g.nz[i]; o[6] = g.area[i];
    }
    return g.n_tri;
}

int emu_n_passbands() { return int(g_passbands.size()); }
void emu_set_star_flux(int on) { g_star_flux = on; }
void emu_set_lx_table(int on) { g_lx_table = on; }
int emu_star_flux() { return g_star_flux; }

long long emu_last_lx_trials() { return g_last_lx_trials; }

// host resolver only: h[Nx], F[Nx] of the initial state + info (no device involved)
int emu_resolve(const fb200_args_t* a, fb200_model_info_t* info, double* h, double* F, char* err, size_t err_len) {
    Resolved res;
    std::string msg;
    const int rc = resolve(*a, res, msg);
    if (rc != FB200_OK) {
        if (err && err_len) std::snprintf(err, err_len, "%s", msg.c_str());
        return rc;
    }
    if (info) *info = res.info;
    if (h) std::memcpy(h, res.h.data(), sizeof(double) * res.h.size());
    if (F) std::memcpy(F, res.F.data(), sizeof(double) * res.F.size());
    return FB200_OK;
}

// policy_trapz (the radial trapezoid of the kernels) under the host policy, on caller data: n <= FB200_NX_MAX points
double emu_trapz(const double* x, const double* y, int n, int first, int last) {
    std::vector<double> xs(FB200_NX_MAX + 1, 0.), ys(FB200_NX_MAX + 1, 0.);
    for (int i = 0; i < n; ++i) { xs[i] = x[i]; ys[i] = y[i]; }
    HostEx ex;
    return policy_trapz(ex, xs, ys, first, last);
}

double emu_planck_nu1_nu2(double T, double nu1, double nu2, double tol, int* trials) { return planck_nu1_nu2(T, nu1, nu2, tol, trials); }

// the same through the table of the band ratio nu2 / nu1 (lx_table.h), as the kernels evaluate it; *walked = rings the table did not answer
void emu_planck_tab(const double* T, long n, double nu1, double nu2, double tol, double* out, long* walked) {
    size_t len = 0;
    const double* tab = lx_table_get(nu2 / nu1, &len);
    long w = 0;
    for (long i = 0; i < n; ++i) {
        int trials = 0;
        out[i] = planck_nu1_nu2_tab(tab, T[i], nu1, nu2, tol, &trials);
        if (tab) {
            const double x1 = FB_H_OVER_KB / T[i] * nu1;
            int t2 = 0;
            if (!(x1 >= FB_LX_X_LO && x1 < FB_LX_X_HI) || (lx_table_q(tab, x1, &t2), t2 < 0)) ++w;
        } else {
            ++w;
        }
    }
    if (walked) *walked = w;
}

}  // extern "C"
