// Builder of the band-integral table (lx_table.h): the breakpoints of the adaptive walk as a function of x1 = h nu1 / kT,
// narrowed down to slivers of 2e-9 relative, and a polynomial fit of q(x1) = J(x1) e^{x1} on every piece in between.
// Host code; runs once per band ratio and process (~50 ms on 8 threads, 2e5 walks), the result is uploaded with every
// ensemble that uses it.  A band ratio whose walk has no tidy piece structure gets no table (the kernels walk).
//
// Reference: Spectrum::Planck_nu1_nu2, cpp/src/spectrum.cpp:26-37 (what the walk restates); planck_walk, disc_engine.h
// (the walk this table reproduces).  Checked against the reference's own function in tests/test_lx_table.py.
#include "lx_table.h"

#include <math.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

#include "disc_engine.h"

namespace fb200 {
namespace {

constexpr int kMaxTrials = 56;
constexpr double kScale = 288230376151711744.0;  // 2^58 ~ the frequency of 1.2 keV: the canonical walk runs at the band's own magnitude,
                                                 // so its comparisons against eps_mach behave as in a real call (scaling by 2^k is exact)
// Where a step's error ratio comes close to the growth limit's floor 5^-5 the walk itself is ill-conditioned: the error estimate
// there is the small remainder of cancelling terms (relative rounding noise ~1e-9), the step factor 0.9 err^(-1/5) hands a fifth of
// that on to every later step, and the result carries ~1e-12 of noise from point to point — in the reference as much as here.
// No polynomial follows noise: a sub-piece is not split below kMinWidth (relative), and a fit within kNoiseTol is accepted there.
constexpr double kMinWidth = 2e-6, kNoiseTol = 2e-11;
constexpr long kMaxWalks = 800000;
constexpr double kMergeGap = 1e-6;   // slivers closer than this (relative) become one
constexpr double kMaxWalkWidth = 1e-2;  // ln-measure of the table's range that may be left to the walk (the range is ln 2^13 = 9)
constexpr double kEdgeWidth = 2e-9;  // relative width of the slivers around the breakpoints (Builder::refine)
  // budget of a build (a normal one takes ~2e5)
constexpr double kTol = 1e-4;                    // eps_rel of planck_nu1_nu2 (FB_KL(CK_TOL))

// What one walk decided: per try_step call a code (bit 4: the step was clipped to the end of the band; low bits: 1 rejected with
// the factor floored at 0.2, 2 rejected, 3 accepted with the growth limit 4.5, 4 accepted and grown, 5 accepted) and the two
// continuous quantities the decisions were taken on: the SIGNED error ratio and (t + dt - nu1) / (nu2 - nu1) before clipping.
struct Trace {
    double x = 0.;
    int n = 0;
    bool overflow = false;
    unsigned char sig[kMaxTrials];
    double err[kMaxTrials], clip[kMaxTrials];
    // Decisions up to and including the accepted step that was clipped to the end of the band.  What may follow it is rounding:
    // t + (nu2 - t) can miss nu2 by an ulp, and the walk then closes with one more step of that length, which adds nothing to the
    // integral but would otherwise make every such ulp pattern a piece of its own.
    int n_eff() const {
        const int m = n < kMaxTrials ? n : kMaxTrials;
        for (int k = 0; k < m; ++k)
            if (sig[k] >= 16 + 3) return k + 1;
        return m;
    }
    bool same_decisions(const Trace& o) const {
        const int m = n_eff();
        return m == o.n_eff() && overflow == o.overflow && std::equal(sig, sig + m, o.sig);
    }
};

// planck_walk (disc_engine.h) statement for statement, recording its decisions.  tests: fb200_lx_table_selftest compares the two.
void traced_walk(double hkT, double nu1, double nu2, double tol, Trace& T) {
    const double b1 = FB_KC(CK_B1), b3 = FB_KC(CK_B3), b4 = FB_KC(CK_B4), b6 = FB_KC(CK_B6);
    const double db1 = FB_KC(CK_DB1), db3 = FB_KC(CK_DB3), db4 = FB_KC(CK_DB4), db5 = FB_KC(CK_DB5), db6 = FB_KC(CK_DB6);
    const double c3 = FB_KC(CK_C3), c4 = FB_KC(CK_C4), c6 = 7. / 8.;
    const double eps_mach = 2.220446049250313e-16;
    const double tiny_w = 5.551115123125783e-17;
    double x = 0., t = nu1, dt = 1e-2 * (nu2 - nu1);
    int fails = 0, n = 0;
    T.overflow = false;
    double wt;
    {
        const double a = hkT * t;
        wt = fb_exp_tab(-((a < 700.) ? a : 700.));
    }
    double k1 = (p3(t) * wt) * fb_rcp(1. - wt);
    while (nu2 - t > eps_mach) {
        unsigned char code = 0;
        const double clip_ratio = ((t + dt) - nu1) / (nu2 - nu1);
        if ((t + dt) - nu2 > eps_mach) {
            dt = nu2 - t;
            code = 16;
        }
        double w3, w4, w5, w6;
        {
            const double u = (hkT * dt) * 0.025;
            const double v = fb_exp_tab(-((u < 700.) ? u : 700.));
            const double v2 = v * v, v3 = v2 * v, v5 = v3 * v2, v6 = v3 * v3, v11 = v6 * v5, v12 = v6 * v6, v24 = v12 * v12;
            const double v35 = v24 * v11;
            w3 = wt * v12; w4 = wt * v24; w6 = wt * v35; w5 = wt * (v35 * v5);
        }
        const bool no_rcp = wt < tiny_w;
        double k3 = p3(t + c3 * dt) * w3, k4 = p3(t + c4 * dt) * w4, k5 = p3(t + dt) * w5, k6 = p3(t + c6 * dt) * w6;
        if (!no_rcp) {
            k3 *= fb_rcp(1. - w3); k4 *= fb_rcp(1. - w4); k5 *= fb_rcp(1. - w5); k6 *= fb_rcp(1. - w6);
        }
        const double xnew = fma(dt, fma(b6, k6, fma(b4, k4, fma(b3, k3, b1 * k1))), x);
        const double xerr = dt * fma(db6, k6, fma(db5, k5, fma(db4, k4, fma(db3, k3, db1 * k1))));
        const double scale = fb_rcp14(tol * (fabs(x) + fabs(dt) * fabs(k1)));
        const double err = fabs(xerr) * scale;
        if (n < kMaxTrials) {
            T.err[n] = xerr * scale;
            T.clip[n] = clip_ratio;
        } else {
            T.overflow = true;
        }
        const bool reject = err > 1.0;
        const double floor_ = FB_KL(CK_FLOOR);
        double fac;
        bool limited = false;
        if (!reject && !(floor_ < err)) {
            fac = 4.5;
            limited = true;
        } else {
            fac = FB_KL(CK_NINE_TENTHS) * fb_pow_step_factor((err < 100.) ? err : 100., reject);
        }
        if (reject) {
            limited = fac < FB_KL(CK_FIFTH);
            dt *= limited ? FB_KL(CK_FIFTH) : fac;
            if (n < kMaxTrials) T.sig[n] = code | (limited ? 1 : 2);
            ++n;
            if (++fails >= 500) break;
            continue;
        }
        fails = 0;
        t += dt;
        if (err < 0.5) dt *= fac;
        if (n < kMaxTrials) T.sig[n] = code | ((err < 0.5) ? (limited ? 3 : 4) : 5);
        ++n;
        x = xnew;
        k1 = k5;
        wt = w5;
    }
    T.x = x;
    T.n = n;
}

uint64_t bits_of(double x) { uint64_t b; memcpy(&b, &x, 8); return b; }
double from_bits(uint64_t b) { double x; memcpy(&x, &b, 8); return x; }

struct Builder {
    double r;
    std::atomic<long> traces{0};
    double c2 = 0.;  // bound on the second derivative (in ln x1) of the decision variables, relative to their thresholds

    void trace(double x1, Trace& T) {
        traces.fetch_add(1, std::memory_order_relaxed);
        traced_walk(x1 / kScale, kScale, r * kScale, kTol, T);
    }
    // q(x1) = J(x1) e^{x1}; *id identifies the decision sequence (try_step count in the low byte, a hash of the codes above it)
    double q_of(double x1, long* id) {
        Trace T;
        trace(x1, T);
        if (id) {
            unsigned long h = 1469598103934665603ul;
            const int m = T.n_eff();
            for (int k = 0; k < m; ++k) h = (h ^ T.sig[k]) * 1099511628211ul;
            *id = T.overflow ? -1 : long(((h >> 16) << 8) | (unsigned long)(m & 255));
        }
        const long double s = 1.0L / ((long double)kScale * kScale * kScale * kScale);
        return double((long double)T.x * s * expl((long double)x1));
    }
    // closest approach of a decision variable to one of its thresholds, relative
    static double margin(const Trace& T) {
        static const double th[4] = {1.0, 0.5, FB_KL(CK_FLOOR), 91.125};
        double m = 1e300;
        const int n = T.n_eff();
        for (int k = 0; k < n; ++k) {
            const double e = fabs(T.err[k]);
            for (double t : th) m = std::min(m, fabs(e - t) / t);
            m = std::min(m, fabs(T.clip[k] - 1.));
        }
        return m;
    }
    // Where the walk's decisions change inside [a, b]: neighbours that decide differently are bisected down to kEdgeWidth
    // (relative), and that last interval is handed back as an EDGE — a sliver in which the table does not answer and the kernels
    // walk.  A decision flips where an error ratio crosses its threshold, and the ratio carries ~1e-10 of rounding noise: inside a
    // band of that width the decisions alternate from one double to the next (in the reference too, with its own roundings), so
    // there is no sharp breakpoint to find; a sliver of a few 1e-9 covers the band, and a ring lands in one about once in 1e8.
    // Where the neighbours agree, a pair of crossings can hide in between only if a decision variable comes within
    // c2 w^2 / 8 of a threshold.
    void refine(double a, const Trace& Ta, double b, const Trace& Tb, std::vector<std::pair<double, double>>& out) {
        const bool same = Ta.same_decisions(Tb);
        const double w = log(b / a);
        if (same) {
            if (std::min(margin(Ta), margin(Tb)) > c2 * w * w * 0.125) return;
            if (w <= kEdgeWidth) return;  // agreeing neighbours this close: nothing in between that an edge would not cover
        } else if (w <= kEdgeWidth || bits_of(b) - bits_of(a) <= 1) {
            out.push_back({a, b});
            return;
        }
        if (traces.load(std::memory_order_relaxed) > kMaxWalks) return;  // a band whose walk has no tidy piece structure: give up (build())
        const double m = from_bits((bits_of(a) + bits_of(b)) / 2);
        Trace Tm;
        trace(m, Tm);
        refine(a, Ta, m, Tm, out);
        refine(m, Tm, b, Tb, out);
    }
};

struct Sub {
    double a, b;  // first and last double of the sub-piece
    double mid, inv_half;
    double c[FB_LX_DEG + 1];
    int trials;
    double err;
};

template <class F>
void parallel_for(int n, int n_threads, F f) {
    if (n_threads <= 1 || n < 2) {
        for (int i = 0; i < n; ++i) f(i);
        return;
    }
    std::atomic<int> next{0};
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t)
        th.emplace_back([&] {
            for (int i; (i = next.fetch_add(1)) < n;) f(i);
        });
    for (auto& t : th) t.join();
}

// the polynomial of lx_table_q, evaluated exactly as the kernels do
double horner(const Sub& s, double x1) {
    const double u = (x1 - s.mid) * s.inv_half;
    double q = s.c[FB_LX_DEG];
    for (int j = FB_LX_DEG - 1; j >= 0; --j) q = fma(q, u, s.c[j]);
    return q;
}

// Fit of one piece [a, b] (doubles, inclusive), split until the polynomial meets `tol` at the check points.  A sub-piece the
// polynomial cannot follow even when narrow (a change of decisions the search did not see) is flagged trials = -1: the kernels
// walk there.
void fit_piece(Builder& B, double a, double b, double tol, std::vector<Sub>& out, double* worst) {
    constexpr int D = FB_LX_DEG, NN = D + 1;
    std::vector<std::pair<double, double>> todo{{a, b}};
    while (!todo.empty()) {
        const double pa = todo.back().first, pb = todo.back().second;
        todo.pop_back();
        Sub s;
        s.a = pa; s.b = pb;
        s.mid = 0.5 * (pa + pb);
        const double half = 0.5 * (pb - pa);
        s.inv_half = (half > 0.) ? 1. / half : 0.;
        long double fv[NN], cheb[NN];
        long id0 = -1;
        bool ok = true;
        for (int j = 0; j < NN; ++j) {
            const double u = cos(M_PI * (j + 0.5) / NN);
            const double x = std::min(std::max(s.mid + half * u, pa), pb);
            long id;
            fv[j] = B.q_of(x, &id);
            if (j == 0) id0 = id;
            ok = ok && id == id0 && id >= 0;
        }
        for (int k = 0; k < NN; ++k) {
            long double sum = 0;
            for (int j = 0; j < NN; ++j) sum += fv[j] * cosl((long double)M_PI * k * (j + 0.5L) / NN);
            cheb[k] = 2.0L * sum / NN;
        }
        cheb[0] *= 0.5L;
        // Chebyshev -> monomial in u (T_0 = 1, T_1 = u, T_{k+1} = 2 u T_k - T_{k-1})
        long double mono[NN] = {0}, Tk0[NN] = {0}, Tk1[NN] = {0}, Tk2[NN];
        Tk0[0] = 1;
        Tk1[1] = 1;
        mono[0] += cheb[0];
        for (int j = 0; j < NN; ++j) mono[j] += cheb[1] * Tk1[j];
        for (int k = 2; k < NN; ++k) {
            for (int j = 0; j < NN; ++j) Tk2[j] = (j > 0 ? 2 * Tk1[j - 1] : 0) - Tk0[j];
            for (int j = 0; j < NN; ++j) {
                mono[j] += cheb[k] * Tk2[j];
                Tk0[j] = Tk1[j];
                Tk1[j] = Tk2[j];
            }
        }
        for (int j = 0; j < NN; ++j) s.c[j] = double(mono[j]);
        s.trials = int(id0 & 255);
        double maxe = 0.;
        constexpr int kChecks = 2 * D + 5;
        for (int j = 0; j < kChecks && ok; ++j) {
            double x;
            if (j == 0) x = pa;
            else if (j == 1) x = pb;
            else x = std::min(std::max(s.mid + half * (-1. + 2. * (j - 2 + 0.37) / (kChecks - 2)), pa), pb);
            long id;
            const double qv = B.q_of(x, &id);
            ok = ok && id == id0;
            maxe = std::max(maxe, fabs(horner(s, x) - qv) / fabs(qv));
        }
        s.err = maxe;
        const bool good = ok && maxe <= tol, noise_only = ok && maxe <= kNoiseTol;
        const double width = (pb - pa) / pa;
        if (!good && width > (noise_only ? kMinWidth : 8. * kEdgeWidth)) {
            const double m = from_bits((bits_of(pa) + bits_of(pb)) / 2);
            todo.push_back({from_bits(bits_of(m) + 1), pb});
            todo.push_back({pa, m});
            continue;
        }
        if (!noise_only) s.trials = -1;
        else if (maxe > *worst) *worst = maxe;
        out.push_back(s);
    }
}

struct Built {
    std::vector<double> blob;
    bool ok = false;
};

void build(double r, Built& out) {
    const bool debug = getenv("FB200_DEBUG") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    Builder B;
    B.r = r;
    int n_threads = int(std::thread::hardware_concurrency());
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 16) n_threads = 16;
    const double lo = FB_LX_X_LO, hi = FB_LX_X_HI;
    constexpr int N0 = 32768;
    std::vector<double> X(N0);
    std::vector<Trace> S(N0);
    for (int i = 0; i < N0; ++i) X[i] = lo * exp(log(hi / lo) * double(i) / (N0 - 1));
    X[0] = lo;
    X[N0 - 1] = hi;
    constexpr int kBlocks = 256, kPer = N0 / kBlocks;
    parallel_for(kBlocks, n_threads, [&](int blk) {
        for (int i = blk * kPer; i < (blk + 1) * kPer; ++i) B.trace(X[i], S[i]);
    });
    for (const Trace& t : S)
        if (t.overflow) return;
    // curvature of the decision variables over triples of samples that decide alike (x 16 for safety)
    {
        const double w = log(X[1] / X[0]);
        double c2 = 0.;
        for (int i = 1; i + 1 < N0; ++i) {
            if (!S[i - 1].same_decisions(S[i]) || !S[i].same_decisions(S[i + 1])) continue;
            for (int k = 0; k < S[i].n_eff(); ++k) {
                const double e = fabs(S[i].err[k]);
                const double ref = std::max(FB_KL(CK_FLOOR), std::min(e, 91.125));
                const double d2 = (S[i + 1].err[k] - 2. * S[i].err[k] + S[i - 1].err[k]) / (w * w) / ref;
                if (std::isfinite(d2)) c2 = std::max(c2, fabs(d2));
                const double d2c = (S[i + 1].clip[k] - 2. * S[i].clip[k] + S[i - 1].clip[k]) / (w * w);
                if (std::isfinite(d2c)) c2 = std::max(c2, fabs(d2c));
            }
        }
        B.c2 = 16. * c2;
    }
    std::vector<std::vector<std::pair<double, double>>> found(kBlocks);
    parallel_for(kBlocks, n_threads, [&](int blk) {
        const int i1 = std::min((blk + 1) * kPer, N0 - 1);
        for (int i = blk * kPer; i < i1; ++i) B.refine(X[i], S[i], X[i + 1], S[i + 1], found[blk]);
    });
    // the slivers in order, touching ones merged; between them the pieces
    std::vector<std::pair<double, double>> slivers;
    for (const auto& f : found)
        for (const auto& iv : f) {
            // touching ones merged, and near ones too: where an error ratio runs along its threshold instead of crossing it, the
            // decisions alternate over a whole stretch — one walked stretch instead of hundreds of slivers
            if (!slivers.empty() && iv.first <= slivers.back().second * (1. + kMergeGap)) slivers.back().second = std::max(slivers.back().second, iv.second);
            else slivers.push_back(iv);
        }
    if (B.traces.load() > kMaxWalks || slivers.size() > size_t(FB_LX_MAX_SUB) / 3) {
        if (debug) fprintf(stderr, "[fb200] Lx table for band ratio %.17g: %zu breakpoints after %ld walks, the kernels will walk\n", r, slivers.size(), B.traces.load());
        return;
    }
    struct Piece { double a, b; bool sliver; };  // first and last double
    std::vector<Piece> pieces;
    {
        double a = lo;
        for (const auto& iv : slivers) {
            // the sliver is (iv.first, iv.second]: iv.first still decides like the piece on its left
            if (iv.first >= a) pieces.push_back({a, iv.first, false});
            pieces.push_back({from_bits(bits_of(iv.first) + 1), iv.second, true});
            a = from_bits(bits_of(iv.second) + 1);
        }
        const double last = from_bits(bits_of(hi) - 1);
        if (a <= last) pieces.push_back({a, last, false});
    }
    const size_t n_pieces = pieces.size();
    const auto t1 = std::chrono::steady_clock::now();

    std::vector<std::vector<Sub>> subs(n_pieces);
    std::vector<double> worst(n_pieces, 0.);
    parallel_for(int(n_pieces), n_threads, [&](int i) {
        if (pieces[i].sliver) {
            Sub s{};
            s.a = pieces[i].a; s.b = pieces[i].b; s.mid = s.a; s.inv_half = 0.; s.trials = -1;
            subs[i].push_back(s);
        } else {
            fit_piece(B, pieces[i].a, pieces[i].b, 2e-12, subs[i], &worst[i]);
        }
    });
    std::vector<Sub> all;
    for (auto& v : subs) {
        std::sort(v.begin(), v.end(), [](const Sub& p, const Sub& q) { return p.a < q.a; });
        all.insert(all.end(), v.begin(), v.end());
    }
    const int n_sub = int(all.size());
    bool tiles = n_sub > 0 && n_sub <= FB_LX_MAX_SUB && all[0].a == lo;
    for (int k = 0; tiles && k + 1 < n_sub; ++k) tiles = bits_of(all[k + 1].a) == bits_of(all[k].b) + 1;  // the sub-pieces must tile [lo, hi)
    double walk_width = 0.;  // ln-measure of the part of the range the kernels walk
    int n_walk = 0;
    for (const Sub& sb : all)
        if (sb.trials < 0) {
            walk_width += log(from_bits(bits_of(sb.b) + 1) / sb.a);
            ++n_walk;
        }
    if (!tiles || walk_width > kMaxWalkWidth) {
        if (debug) fprintf(stderr, "[fb200] Lx table for band ratio %.17g: %d sub-pieces, %d of them (ln-width %.3g) without a fit, the kernels will walk\n", r, n_sub, n_walk, walk_width);
        return;
    }
    double worst_all = 0.;
    for (double w : worst) worst_all = std::max(worst_all, w);

    out.blob.assign(lx_table_doubles(n_sub), 0.);
    double* tb = out.blob.data();
    tb[0] = r; tb[1] = double(n_sub); tb[2] = kTol; tb[3] = worst_all; tb[4] = double(slivers.size()); tb[5] = double(n_walk); tb[6] = walk_width;
    for (int k = 0; k < n_sub; ++k) {
        tb[FB_LX_O_EDGE + k] = all[k].a;
        double* rec = tb + FB_LX_O_REC + FB_LX_REC * k;
        rec[0] = all[k].mid;
        rec[1] = all[k].inv_half;
        for (int j = 0; j <= FB_LX_DEG; ++j) rec[2 + j] = all[k].c[j];
        rec[14] = double(all[k].trials);
    }
    for (int k = n_sub; k <= FB_LX_MAX_SUB; ++k) tb[FB_LX_O_EDGE + k] = hi;
    int* cells = reinterpret_cast<int*>(tb + FB_LX_O_CELL);
    auto sub_of = [&](double x) { return int(std::upper_bound(tb + FB_LX_O_EDGE, tb + FB_LX_O_EDGE + n_sub, x) - (tb + FB_LX_O_EDGE)) - 1; };
    for (int c = 0; c < FB_LX_CELLS; ++c) {
        const double xl = from_bits(uint64_t(uint32_t(c + FB_LX_CELL0) << 14) << 32);
        const double xr = from_bits(uint64_t(uint32_t(c + 1 + FB_LX_CELL0) << 14) << 32);
        cells[2 * c] = sub_of(xl);
        cells[2 * c + 1] = sub_of(from_bits(bits_of(xr) - 1));
    }
    out.ok = true;
    if (debug) {
        const auto t2 = std::chrono::steady_clock::now();
        fprintf(stderr, "[fb200] Lx table for band ratio %.6g: %zu breakpoints, %d sub-pieces (%d walked, ln-width %.2g), fit error <= %.2g, %ld walks, %.1f + %.1f ms on %d threads\n",
                r, slivers.size(), n_sub, n_walk, walk_width, worst_all, B.traces.load(), std::chrono::duration<double, std::milli>(t1 - t0).count(),
                std::chrono::duration<double, std::milli>(t2 - t1).count(), n_threads);
    }
}

std::mutex g_mutex;
std::map<uint64_t, Built*> g_tables;

}  // namespace

const double* lx_table_get(double r, size_t* doubles) {
    if (doubles) *doubles = 0;
    if (!(r > 1.0) || !(r < 1e6)) return nullptr;
    if (const char* c = getenv("FB200_LX_TABLE"))
        if (c[0] == '0') return nullptr;
    std::lock_guard<std::mutex> g(g_mutex);
    Built*& b = g_tables[bits_of(r)];
    if (!b) {
        b = new Built;  // lives as long as the process: ensembles keep pointers into it while they upload
        build(r, *b);
    }
    if (!b->ok) return nullptr;
    if (doubles) *doubles = b->blob.size();
    return b->blob.data();
}

}  // namespace fb200

// Diagnostic entry point (include/freddi_b200.h): the table against the direct walk at `n` temperatures, log-uniform in
// x1 = h nu1 / kT over [x_lo, x_hi], for the band [nu1, nu2] (Hz).  Host arithmetic only.
extern "C" int fb200_lx_table_selftest(double nu1, double nu2, double x_lo, double x_hi, long n, unsigned long long seed,
                                       double* max_rel_err, long* trials_mismatches, double* worst_x1, int* n_sub) {
    using namespace fb200;
    size_t len = 0;
    const double r = nu2 / nu1;
    const double* tb = lx_table_get(r, &len);
    if (!tb) return FB200_ERR_UNSUPPORTED;
    if (n_sub) *n_sub = int(tb[1]);
    double worst = 0., wx = 0.;
    long mism = 0;
    uint64_t s = seed ? seed : 88172645463325252ull;
    for (long i = 0; i < n; ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        const double u = double(s >> 11) * (1.0 / 9007199254740992.0);
        const double x1 = x_lo * exp(u * log(x_hi / x_lo));
        if (!(x1 >= FB_LX_X_LO) || !(x1 < FB_LX_X_HI)) continue;
        const double hkT = x1 / nu1;
        int n_walk = 0, n_tab = 0;
        const double direct = planck_walk(hkT, nu1, nu2, kTol, &n_walk);
        const double xx = hkT * nu1;  // what the kernels look up
        if (!(xx >= FB_LX_X_LO && xx < FB_LX_X_HI)) continue;
        const double q = lx_table_q(tb, xx, &n_tab);
        if (n_tab < 0) continue;  // a sliver at a breakpoint: the kernels walk there
        const double tab = p4(nu1) * (q * fb_exp_tab(-xx));
        const double e = fabs(tab - direct) / fabs(direct);
        if (!(e <= worst)) { worst = e; wx = x1; }
        if (n_walk != n_tab) ++mism;
    }
    if (max_rel_err) *max_rel_err = worst;
    if (trials_mismatches) *trials_mismatches = mism;
    if (worst_x1) *worst_x1 = wx;
    return FB200_OK;
}
