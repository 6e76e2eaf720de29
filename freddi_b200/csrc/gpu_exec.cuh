// CUDA execution policy of DiscEngine: one CTA of kNT threads per model, kPPT = 1024 / kNT radial points per
// thread (point i = tid + k*kNT lives in register slot k).  Default 256 threads x 4 points, two CTAs per SM
// (128 registers per thread, no spills; ~110 KB of shared memory per CTA) — measured best of {1024x1, 512x2,
// 512x1, 256x2} (profiles/r01_layout_variants.txt).  Reductions are warp-shuffle trees finished through shared memory; every
// reduction returns its result to all threads so the orchestration code stays uniform across the CTA.
#pragma once
#include <cuda_runtime.h>

#include "evolve.cuh"

namespace fb200 {

#ifndef FB_NT
#define FB_NT 256
#endif
#ifndef FB_CTAS_PER_SM
#define FB_CTAS_PER_SM 2
#endif
// unroll factors of the per-thread slot loops (ILP against kernel size; measured in profiles/r01_layout_variants.txt)
#ifndef FB_UNROLL_POINTS
#define FB_UNROLL_POINTS 1
#endif
#ifndef FB_UNROLL_MAX
#define FB_UNROLL_MAX 1
#endif
#ifndef FB_UNROLL_SUMS
#define FB_UNROLL_SUMS 1
#endif
#define FB_PRAGMA(x) _Pragma(#x)
#define FB_UNROLL(n) FB_PRAGMA(unroll n)
constexpr int kNT = FB_NT;                 // threads per CTA
constexpr int kPPT = FB200_NX_MAX / kNT;   // radial points per thread (FB_BIG builds: the upper bound, see GpuEx::n_slots)
constexpr int kNW = kNT / 32;              // warps per CTA
constexpr int kCtasPerSm = FB_CTAS_PER_SM;
static_assert(kNT * kPPT == FB200_NX_MAX && kNW >= 1 && kNW <= 32, "FB_NT must divide FB200_NX_MAX");
// FB_BIG: the large-grid build (Nx up to FB200_NX_MAX = 16384).  Same engine, same kernels; the per-point arrays do not
// fit into shared memory (22 arrays x 128 KB), so they all live in the CTA's global scratch slot (L2-resident for the
// small ensembles such grids are used with) and every per-thread slot loop has a run-time trip count ceil(Nx / kNT).
// Shared memory keeps the exp table, the reduction scratch, the model constants and the output configuration.

// The CTA's dynamic shared memory.  Every kernel of this library declares the same window; all array
// accessors below index it with compile-time offsets, so the compiler emits LDS/STS with immediate offsets.
// (declared in fb200_math.h; its first 64 doubles hold the 2^(j/32) table (values, then tails) of fb_exp_tab / fb_pow_pos)
__constant__ double c_exp2_table[64] = {FB_EXP2_TABLE, FB_EXP2_TABLE_LO};

// Shared memory per CTA: reduced-system PCR ping-pong, F, hn, s_cc, s_frac (plain index), s_l, s_u, s_d, s_rhs,
// p_a, p_c, p_g (padded index), then the reduction scratch, the model constants and the output configuration.
// Global scratch slot per CTA: R, ca, cb, cc, frac, c1, hm175, hotR, Gb, sgA, qB, tv4 (12 point arrays, read once per step).
struct SmemLayout {
    static constexpr int N = FB200_NX_MAX, NP = FB200_NPAD, NR = FB200_NRED_MAX;
    static constexpr int o_exptab = 0;  // fixed: fb200_math.h reads the table at fb_smem[0..63]
    static constexpr int o_lu0 = 64, o_lu1 = o_lu0 + 2 * NR, o_rr0 = o_lu1 + 2 * NR, o_rr1 = o_rr0 + NR;
    static constexpr int o_F = o_rr1 + NR, o_hn = o_F + N, o_cc = o_hn + N, o_frac = o_cc + N, o_K = o_frac + N;
    static constexpr int o_l = o_K + N, o_u = o_l + NP, o_d = o_u + NP, o_rhs = o_d + NP;
    static constexpr int o_pa = o_rhs + NP, o_pc = o_pa + NP, o_pg = o_pc + NP;
#ifdef FB_BIG
    static constexpr int kPointDoubles = 64;  // only the exp table; o_lu0..o_pg above are unused
#else
    static constexpr int kPointDoubles = o_pg + NP;
#endif
    // reduction scratch
#ifdef FB_BIG
    static constexpr int o_redA = kPointDoubles, o_redQ = o_redA + 2 * 32, o_sums = o_redQ + kNW * FB200_MAXQ;
#else
    // The cross-warp partial sums of reduce_sums / reduce_sums3 ([kNW][FB200_MAXQ]) alias the first PCR ping-pong buffer:
    // sums are only taken in the row phase and in the read-out kernels, when the reduced system is dead (the next solve
    // rebuilds it from scratch).  This keeps the CTA at <= 113 KB so that two CTAs fit into the SM's 228 KB.
    static constexpr int o_redA = kPointDoubles, o_redQ = o_lu0, o_sums = o_redA + 2 * 32;
    static_assert(kNW * FB200_MAXQ <= 2 * NR, "reduce_sums scratch must fit into the first PCR buffer");
#endif
    // o_count: FB_CNT_N bookkeeping counters (long long) of the current work item, then the wind-update record (Mdot, first|last)
    static constexpr int o_scal = o_sums + FB200_MAXQ, o_count = o_scal + 8, o_lampref = o_count + FB_CNT_N + 2, o_starsum = o_lampref + FB200_MAX_LAMBDAS;
    static constexpr int o_model = o_starsum + 3 * (FB200_MAX_LAMBDAS + FB200_MAX_PASSBANDS);
    static constexpr int kModelDoubles = int((sizeof(fb200_model_dev_t) + 7) / 8);
    static constexpr int o_cfg = o_model + kModelDoubles;
    static constexpr int kCfgDoubles = int((sizeof(OutCfg) + 7) / 8);
    static constexpr int kTotalDoubles = o_cfg + kCfgDoubles + 2;
    static constexpr size_t bytes() { return sizeof(double) * size_t(kTotalDoubles); }
#ifndef FB_BIG
    // B200: 233,472 B of shared memory per SM, 1 KB reserved per CTA -> two resident CTAs need <= 115,712 B each.
    // (One CTA per SM costs 38 % of the throughput: measured 59 ms -> 95 ms per launch of the bench.)
    static_assert(FB_CTAS_PER_SM != 2 || sizeof(double) * size_t(kTotalDoubles) <= 115712, "shared memory per CTA too large for 2 CTAs/SM");
#endif
#ifdef FB_BIG
    // R ca cb cc frac c1 hm175 hotR Gb sgA qB tv4 | F hn s_cc s_frac s_K | s_l s_u s_d s_rhs p_a p_c p_g | lu0 lu1 (2 NR each) rr0 rr1
    static constexpr int kScratchArrays = 24;
    static constexpr size_t kScratchDoublesPerCta = size_t(kScratchArrays) * FB200_NX_MAX + 6 * size_t(NR);
#else
    static constexpr int kScratchArrays = 12;
    static constexpr size_t kScratchDoublesPerCta = size_t(kScratchArrays) * FB200_NX_MAX;
#endif
    static constexpr size_t scratch_doubles_per_cta() { return kScratchDoublesPerCta; }
};

template <int OFF> struct ShD {   // window of doubles at a compile-time offset of the CTA's shared memory
    __device__ __forceinline__ double& operator[](int i) const { return fb_smem[OFF + i]; }
};
struct ShDyn {                    // same with a run-time offset (the ping-pong buffers)
    int off;
    __device__ __forceinline__ double& operator[](int i) const { return fb_smem[off + i]; }
};
struct ShDyn2 {
    int off;
    __device__ __forceinline__ double2& operator[](int i) const { return reinterpret_cast<double2*>(fb_smem + off)[i]; }
};

#ifdef FB_BIG
typedef PtrArrays GpuArrays;  // every array a plain pointer into the CTA's global scratch slot

__device__ __forceinline__ GpuArrays carve_arrays(double* gscratch, const double* h_global) {
    GpuArrays S;
    const size_t N = FB200_NX_MAX, NR = FB200_NRED_MAX;
    double* p = gscratch;
    auto take = [&](size_t n) { double* q = p; p += n; return q; };
    S.h = h_global;
    S.R = take(N); S.ca = take(N); S.cb = take(N); S.cc = take(N); S.frac = take(N);
    S.c1 = take(N); S.hm175 = take(N); S.hotR = take(N); S.Gb = take(N);
    S.sgA = take(N); S.qB = take(N); S.tv4 = take(N);
    S.F = take(N); S.hn = take(N); S.s_cc = take(N); S.s_frac = take(N); S.s_K = take(N);
    S.s_l = take(N); S.s_u = take(N); S.s_d = take(N); S.s_rhs = take(N); S.p_a = take(N); S.p_c = take(N); S.p_g = take(N);
    S.lu[0] = reinterpret_cast<double2*>(take(2 * NR)); S.lu[1] = reinterpret_cast<double2*>(take(2 * NR));
    S.rr[0] = take(NR); S.rr[1] = take(NR);
    S.Tph = S.s_l; S.Tirr = S.s_u; S.Sigma = S.s_d; S.Tvis = S.s_rhs; S.TphX = S.p_a; S.aux = S.p_c; S.wq = S.p_g;
    return S;
}

#else
struct GpuArrays {
    typedef SmemLayout L;
    const double* h;                                   // the ensemble's grid array of this model (global)
    double *R, *ca, *cb, *cc, *frac, *c1, *hm175, *hotR, *Gb, *sgA, *qB, *tv4;  // the CTA's global scratch slot
    ShD<L::o_F> F; ShD<L::o_hn> hn; ShD<L::o_cc> s_cc; ShD<L::o_frac> s_frac; ShD<L::o_K> s_K;
    ShD<L::o_l> s_l; ShD<L::o_u> s_u; ShD<L::o_d> s_d; ShD<L::o_rhs> s_rhs;
    ShD<L::o_pa> p_a; ShD<L::o_pc> p_c; ShD<L::o_pg> p_g;
    // ping-pong pairs: operator[](cur) selects the window arithmetically (a real array indexed at run time
    // would force the whole struct, and the engine object holding it, into local memory)
    struct PP2 { __device__ __forceinline__ ShDyn2 operator[](int cur) const { return ShDyn2{cur ? L::o_lu1 : L::o_lu0}; } } lu;
    struct PP1 { __device__ __forceinline__ ShDyn operator[](int cur) const { return ShDyn{cur ? L::o_rr1 : L::o_rr0}; } } rr;
    // row temporaries alias the solver arrays (rebuilt at the start of every step)
    ShD<L::o_l> Tph; ShD<L::o_u> Tirr; ShD<L::o_d> Sigma; ShD<L::o_rhs> Tvis; ShD<L::o_pa> TphX; ShD<L::o_pc> aux; ShD<L::o_pg> wq;
};

__device__ __forceinline__ GpuArrays carve_arrays(double* gscratch, const double* h_global) {
    GpuArrays S;
    const int N = FB200_NX_MAX;
    S.h = h_global;
    S.R = gscratch + 0 * N; S.ca = gscratch + 1 * N; S.cb = gscratch + 2 * N; S.cc = gscratch + 3 * N; S.frac = gscratch + 4 * N;
    S.c1 = gscratch + 5 * N; S.hm175 = gscratch + 6 * N; S.hotR = gscratch + 7 * N; S.Gb = gscratch + 8 * N;
    S.sgA = gscratch + 9 * N; S.qB = gscratch + 10 * N; S.tv4 = gscratch + 11 * N;
    return S;
}

#endif

__device__ __forceinline__ fb200_model_dev_t* smem_model() { return reinterpret_cast<fb200_model_dev_t*>(fb_smem + SmemLayout::o_model); }
__device__ __forceinline__ OutCfg* smem_cfg() { return reinterpret_cast<OutCfg*>(fb_smem + SmemLayout::o_cfg); }

__device__ __forceinline__ int& fb_smem_int(int double_slot) { return *reinterpret_cast<int*>(fb_smem + double_slot); }
__device__ __forceinline__ unsigned long long fb_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

struct GpuEx {
    typedef GpuArrays Arrays;
    typedef SmemLayout L;
    int tid, lane, warp;
    int flip;
    // per-model global arrays (already offset to this model), null when absent
    const double *gA, *gB, *gCs, *gFm, *gdFm;
    double* gStarT;  // per-triangle effective temperature of the companion star (the CTA's slice of EnsembleDev::star_T)

#ifdef FB_BIG
    int n_slots;  // ceil(Nx / kNT)
    __device__ __forceinline__ void set_nx(int Nx) { n_slots = (Nx + kNT - 1) / kNT; }
    __device__ __forceinline__ int slots() const { return n_slots; }
#else
    __device__ __forceinline__ void set_nx(int) {}
    __device__ __forceinline__ static constexpr int slots() { return kPPT; }
#endif
    __device__ __forceinline__ void init() {
        tid = threadIdx.x; lane = tid & 31; warp = tid >> 5; flip = 0;
        gA = gB = gCs = gFm = gdFm = nullptr;
        gStarT = nullptr;
        set_nx(FB200_NX_MAX);
    }
    // f(i, k): point index i, register slot k (a compile-time constant after unrolling)
    // The slot loops are deliberately NOT unrolled: the kernel is large and unrolling every per-point lambda
    // kPPT times costs more in instruction-cache misses than it gains in ILP (profiles/r01_v3_ncu_full_metrics.txt).
    template <class F> __device__ __forceinline__ void points(F&& f) {
        FB_UNROLL(FB_UNROLL_POINTS)
        for (int k = 0; k < slots(); ++k) f(tid + k * kNT, k);
    }
    // the same with the slot loop fully unrolled: for branch-free lambdas whose per-ring dependency chains should interleave
    template <class F> __device__ __forceinline__ void points_ilp(F&& f) {
#ifdef FB_BIG
        for (int k = 0; k < slots(); ++k) f(tid + k * kNT, k);
#else
#pragma unroll
        for (int k = 0; k < kPPT; ++k) f(tid + k * kNT, k);
#endif
    }
    // one thread per block / per reduced unknown of the tridiagonal solve
    template <class F> __device__ __forceinline__ void blocks(int n, F&& f) {
#ifdef FB_BIG
        for (int b = tid; b < n; b += kNT) f(b);
#else
        if (tid < n) f(tid);
#endif
    }
    template <class F> __device__ __forceinline__ void reduced(int n, F&& f) {
        for (int q = tid; q < n; q += kNT) f(q);
    }
    __device__ __forceinline__ void sync() { __syncthreads(); }
    // hybrid reduced-system solver of DiscEngine::solve_step: needs one thread per reduced unknown, 8 warps
#ifdef FB_BIG
    static constexpr bool kWarpPcr = false;
#else
    static constexpr bool kWarpPcr = (kNT == 256 && FB200_NRED_MAX == 256);
#endif
    __device__ __forceinline__ int thread() const { return tid; }
    __device__ __forceinline__ double shfl_up(double v, int d) const { return __shfl_up_sync(0xffffffffu, v, d); }
    __device__ __forceinline__ double shfl_down(double v, int d) const { return __shfl_down_sync(0xffffffffu, v, d); }
#ifdef FB_PROFILE
    // cycle accounting per phase (thread 0 only): phase p collects the cycles since the previous tick
    long long prof_last;
    long long prof[16];
    __device__ __forceinline__ void tick(int p) {
        if (tid == 0) { const long long c = clock64(); prof[p] += c - prof_last; prof_last = c; }
    }
#else
    __device__ __forceinline__ void tick(int) {}
#endif
    __device__ __forceinline__ double& scalar(int k) { return fb_smem[L::o_scal + k]; }
    // bookkeeping (DiscEngine, FB_CNT_*): thread 0 owns the counters of the current work item in shared memory; only the rare
    // pow fall-backs of the incremental K update are counted by every thread (atomically)
    __device__ __forceinline__ static long long* counters() { return reinterpret_cast<long long*>(fb_smem + L::o_count); }
#ifndef FB_COUNT_MASK  // measurement variants: which counters are compiled in (bit c = FB_CNT_* c)
#define FB_COUNT_MASK 0xff
#endif
    __device__ __forceinline__ void count(int c, long long v) { if (((FB_COUNT_MASK) >> c) & 1) { if (tid == 0) counters()[c] += v; } }
    __device__ __forceinline__ void note_pow_fallbacks(int n) {
        // a 32-bit shared-memory atomic (native ATOMS.ADD; the 64-bit one is a compare-and-swap loop whose registers cost the
        // Picard loop 7 % — measured, profiles/r02_counter_cost.txt); the slot's upper word stays 0 (< 2^31 rings per work item)
        if (((FB_COUNT_MASK) >> FB_CNT_POW_FALLBACKS) & 1)
            if (n) atomicAdd(reinterpret_cast<int*>(counters() + FB_CNT_POW_FALLBACKS), n);
    }
    // per-thread counts of the Lx quadrature's try_step calls: one REDUX per warp and a 32-bit shared-memory atomic (< 2^31 per work item)
    __device__ __forceinline__ void note_lx_trials(int n) {
        if (((FB_COUNT_MASK) >> FB_CNT_LX_TRIALS) & 1) {
            n = __reduce_add_sync(0xffffffffu, n);
            if (lane == 0 && n) atomicAdd(reinterpret_cast<int*>(counters() + FB_CNT_LX_TRIALS), n);
        }
    }
    // CTA-wide OR of a predicate; a full barrier like sync()
    __device__ __forceinline__ bool any_of(bool p) { return __syncthreads_or(p ? 1 : 0) != 0; }
    __device__ __forceinline__ static WindRecord& wind_update() { return *reinterpret_cast<WindRecord*>(fb_smem + L::o_count + FB_CNT_N); }
    __device__ __forceinline__ void note_wind_update(double Mdot, int first, int last) {
        if (tid == 0) wind_update() = WindRecord{Mdot, first, last};
    }
    __device__ __forceinline__ double sum(int q) const { return fb_smem[L::o_sums + q]; }
    __device__ __forceinline__ double lambda_pref(int k) const { return fb_smem[L::o_lampref + k]; }
    __device__ __forceinline__ double& star_T(int t) const { return gStarT[t]; }
    __device__ __forceinline__ double& star_sum(int k) const { return fb_smem[L::o_starsum + k]; }
    template <class F> __device__ __forceinline__ void tri_points(int n, F&& f) {
        for (int t = tid; t < n; t += kNT) f(t);
    }
    // sums over t < n of the three values f(t, a, b, c) produces -> sum(0), sum(1), sum(2)
    template <class F> __device__ __forceinline__ void reduce_sums3(int n, F&& f) {
        double v0 = 0., v1 = 0., v2 = 0.;
        for (int t = tid; t < n; t += kNT) {
            double a, b, c;
            f(t, a, b, c);
            v0 += a; v1 += b; v2 += c;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v0 += __shfl_xor_sync(0xffffffffu, v0, o);
            v1 += __shfl_xor_sync(0xffffffffu, v1, o);
            v2 += __shfl_xor_sync(0xffffffffu, v2, o);
        }
        if (lane == 0) {
            fb_smem[L::o_redQ + warp * FB200_MAXQ + 0] = v0;
            fb_smem[L::o_redQ + warp * FB200_MAXQ + 1] = v1;
            fb_smem[L::o_redQ + warp * FB200_MAXQ + 2] = v2;
        }
        __syncthreads();
        if (tid < 3) {
            double s = 0.;
#pragma unroll
            for (int w = 0; w < kNW; ++w) s += fb_smem[L::o_redQ + w * FB200_MAXQ + tid];
            fb_smem[L::o_sums + tid] = s;
        }
        __syncthreads();
    }
    __device__ __forceinline__ double windA(int i) const { return gA ? __ldg(gA + i) : 0.; }
    __device__ __forceinline__ double windB(int i) const { return gB ? __ldg(gB + i) : 0.; }
    __device__ __forceinline__ double windC_static(int i) const { return gCs ? __ldg(gCs + i) : 0.; }
    __device__ __forceinline__ bool has_windC() const { return gCs != nullptr; }
    __device__ __forceinline__ double Fmagn(int i) const { return gFm ? __ldg(gFm + i) : 0.; }
    __device__ __forceinline__ double dFmagn_dh(int i) const { return gdFm ? __ldg(gdFm + i) : 0.; }

    // One barrier per reduction: consecutive reductions alternate between two scratch rows, so a row is
    // rewritten only after every thread has passed the barrier of the reduction in between.
    template <class F> __device__ __forceinline__ double reduce_max(F&& f) {
        double v = -INFINITY;
        FB_UNROLL(FB_UNROLL_MAX)
        for (int k = 0; k < slots(); ++k) {
            const double x = f(tid + k * kNT, k);
            v = fmax(v, x);  // fmax skips a NaN like `if (x > max)` does
        }
        return reduce_max_of(v);
    }
    // CTA maximum of one value per thread (NaN-skipping), returned to every thread
    __device__ __forceinline__ double reduce_max_of(double v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        double* row = fb_smem + L::o_redA + flip * 32;
        flip ^= 1;
        if (lane == 0) row[warp] = v;
        __syncthreads();
        double r = (lane < kNW) ? row[lane] : -INFINITY;
#pragma unroll
        for (int o = kNW / 2; o > 0; o >>= 1) r = fmax(r, __shfl_xor_sync(0xffffffffu, r, o));
        return __shfl_sync(0xffffffffu, r, 0);
    }
    template <class F> __device__ __forceinline__ int reduce_max_int(F&& f) {
        int v = 0x80000000;
#pragma unroll 1
        for (int k = 0; k < slots(); ++k) v = max(v, f(tid + k * kNT, k));
        v = __reduce_max_sync(0xffffffffu, v);
        int* row = reinterpret_cast<int*>(fb_smem + L::o_redA + flip * 32);
        flip ^= 1;
        if (lane == 0) row[warp] = v;
        __syncthreads();
        return __reduce_max_sync(0xffffffffu, (lane < kNW) ? row[lane] : int(0x80000000));
    }
    template <class F> __device__ __forceinline__ int reduce_min_int(F&& f) {
        int v = 0x7fffffff;
#pragma unroll 1
        for (int k = 0; k < slots(); ++k) v = min(v, f(tid + k * kNT, k));
        v = __reduce_min_sync(0xffffffffu, v);
        int* row = reinterpret_cast<int*>(fb_smem + L::o_redA + flip * 32);
        flip ^= 1;
        if (lane == 0) row[warp] = v;
        __syncthreads();
        return __reduce_min_sync(0xffffffffu, (lane < kNW) ? row[lane] : 0x7fffffff);
    }
    // sums[q] = sum over all points of f(i, k, q), q < nq
    template <class F> __device__ __forceinline__ void reduce_sums(int nq, F&& f) {
        for (int q = 0; q < nq; ++q) {
            double v = 0.;
            FB_UNROLL(FB_UNROLL_SUMS)
            for (int k = 0; k < slots(); ++k) v += f(tid + k * kNT, k, q);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) fb_smem[L::o_redQ + warp * FB200_MAXQ + q] = v;
        }
        __syncthreads();
        if (tid < nq) {
            double s = 0.;
#pragma unroll
            for (int w = 0; w < kNW; ++w) s += fb_smem[L::o_redQ + w * FB200_MAXQ + tid];
            fb_smem[L::o_sums + tid] = s;
        }
        __syncthreads();
    }
};

// Copies the output configuration into shared memory with compile-time indices only: indexing a kernel
// parameter dynamically would make the compiler spill the whole parameter block to local memory.
__device__ __forceinline__ const OutCfg& stage_cfg(const EnsembleDev& E) {
    OutCfg* c = smem_cfg();
    for (int j = threadIdx.x; j < 64; j += kNT) fb_smem[SmemLayout::o_exptab + j] = c_exp2_table[j];
    if (threadIdx.x == 0) {
        c->n_lambdas = E.cfg.n_lambdas; c->cold_disk = E.cfg.cold_disk; c->compute_lx = E.cfg.compute_lx;
        c->record_F = E.cfg.record_F; c->n_cols = E.cfg.n_cols; c->n_rows = E.cfg.n_rows; c->lx_prune = E.cfg.lx_prune;
        c->n_passbands = E.cfg.n_passbands; c->pb_a = E.cfg.pb_a; c->pb_w = E.cfg.pb_w;
        c->star_flux = E.cfg.star_flux; c->star_data = E.cfg.star_data; c->lx_table = E.cfg.lx_table;
#pragma unroll
        for (int k = 0; k < FB200_MAX_PASSBANDS + 2; ++k) c->pb_off[k] = E.cfg.pb_off[k];
#pragma unroll
        for (int k = 0; k < FB200_MAX_PASSBANDS; ++k) c->pb_tdnu[k] = E.cfg.pb_tdnu[k];
#pragma unroll
        for (int k = 0; k < FB200_MAX_LAMBDAS; ++k) {
            const double lam = E.cfg.lambdas[k];
            c->lambdas[k] = lam;
            fb_smem[SmemLayout::o_lampref + k] = FB_DOUBLE_H_C2 / p5(lam);
        }
    }
    __syncthreads();
    return *c;
}

// Loads one model's constants and state into shared memory and binds its global arrays.  All threads call it.
__device__ __forceinline__ void load_model(const EnsembleDev& E, int model, const GpuArrays& S, GpuEx& ex, const double* F_src) {
    fb200_model_dev_t* Msh = smem_model();
    const int tid = threadIdx.x;
    {
        const double* src = reinterpret_cast<const double*>(E.models + model);
        double* dst = reinterpret_cast<double*>(Msh);
        for (int k = tid; k < int(sizeof(fb200_model_dev_t) / sizeof(double)); k += kNT) dst[k] = src[k];
    }
    const int Nx = E.Nx;
    const size_t off = size_t(model) * Nx;
    ex.set_nx(Nx);
#ifndef FB_BIG
#pragma unroll
#endif
    for (int k = 0; k < ex.slots(); ++k) {
        const int i = tid + k * kNT;
        S.F[i] = (i < Nx) ? __ldcg(F_src + i) : 0.;  // L2: another SM may have written it in the previous round
        if (i >= Nx) S.R[i] = 0.;
    }
    // (ex.lampref is filled once per kernel by stage_cfg)
    ex.gA = E.wA ? E.wA + off : nullptr;
    ex.gB = E.wB ? E.wB + off : nullptr;
    ex.gCs = E.wCs ? E.wCs + off : nullptr;
    ex.gFm = E.Fmagn ? E.Fmagn + off : nullptr;
    ex.gdFm = E.dFmagn_dh ? E.dFmagn_dh + off : nullptr;
    ex.gStarT = E.star_T ? E.star_T + size_t(blockIdx.x) * E.star_ntri_max : nullptr;
    __syncthreads();
}

}  // namespace fb200
