// CUDA execution policy of DiscEngine: one CTA per model, one radial point per thread (NT = 1024 threads).
// Reductions are warp-shuffle trees finished through shared memory; every reduction returns its result
// to all threads so the orchestration code stays uniform across the CTA.
#pragma once
#include <cuda_runtime.h>

#include "evolve.cuh"

namespace fb200 {

constexpr int kNT = 1024;      // threads per CTA = FB200_NX_MAX
constexpr int kNW = kNT / 32;  // warps per CTA

struct SmemLayout {
    // 18 point arrays of FB200_NX_MAX doubles, then the small scratch areas
    static constexpr int kPointArrays = 18;
    static constexpr size_t kRedDoubles = 2 * kNW + kNW * FB200_MAXQ + FB200_MAXQ + 8 + FB200_MAX_LAMBDAS;
    static constexpr size_t bytes() {
        return sizeof(double) * (size_t(kPointArrays) * FB200_NX_MAX + kRedDoubles) + sizeof(fb200_model_dev_t) + 64;
    }
};

__device__ __forceinline__ SharedArrays carve_shared(double* base) {
    SharedArrays S;
    const int N = FB200_NX_MAX;
    S.h = base + 0 * N; S.R = base + 1 * N; S.F = base + 2 * N; S.hn = base + 3 * N;
    S.ca = base + 4 * N; S.cb = base + 5 * N; S.cc = base + 6 * N; S.frac = base + 7 * N;
    S.c1 = base + 8 * N; S.hm175 = base + 9 * N; S.hotR = base + 10 * N; S.Gb = base + 11 * N;
    S.lu[0] = reinterpret_cast<double2*>(base + 12 * N);  // 2N doubles
    S.lu[1] = reinterpret_cast<double2*>(base + 14 * N);  // 2N doubles
    S.rr[0] = base + 16 * N;
    S.rr[1] = base + 17 * N;
    // row temporaries alias the PCR buffers
    S.Tph = base + 12 * N; S.Tirr = base + 13 * N; S.Sigma = base + 14 * N; S.Tvis = base + 15 * N;
    S.TphX = base + 16 * N; S.aux = base + 17 * N;
    return S;
}

struct GpuEx {
    int tid, lane, warp;
    StepRegs sr;
    double* redA;     // [2][kNW] double-buffered scalar reductions
    double* redQ;     // [kNW][FB200_MAXQ]
    double* sums;     // [FB200_MAXQ]
    double* scal;     // [8]
    double* lampref;  // [FB200_MAX_LAMBDAS]
    int flip;
    // per-model global arrays (already offset to this model), null when absent
    const double *gA, *gB, *gCs, *gFm, *gdFm;

    __device__ __forceinline__ void init(double* scratch) {
        tid = threadIdx.x; lane = tid & 31; warp = tid >> 5; flip = 0;
        redA = scratch; redQ = redA + 2 * kNW; sums = redQ + kNW * FB200_MAXQ; scal = sums + FB200_MAXQ; lampref = scal + 8;
        gA = gB = gCs = gFm = gdFm = nullptr;
    }
    template <class F> __device__ __forceinline__ void points(F&& f) { f(tid); }
    __device__ __forceinline__ void sync() { __syncthreads(); }
    __device__ __forceinline__ StepRegs& step(int) { return sr; }
    __device__ __forceinline__ double& scalar(int k) { return scal[k]; }
    __device__ __forceinline__ double sum(int q) const { return sums[q]; }
    __device__ __forceinline__ double lambda_pref(int k) const { return lampref[k]; }
    __device__ __forceinline__ double windA(int i) const { return gA ? __ldg(gA + i) : 0.; }
    __device__ __forceinline__ double windB(int i) const { return gB ? __ldg(gB + i) : 0.; }
    __device__ __forceinline__ double windC_static(int i) const { return gCs ? __ldg(gCs + i) : 0.; }
    __device__ __forceinline__ double Fmagn(int i) const { return gFm ? __ldg(gFm + i) : 0.; }
    __device__ __forceinline__ double dFmagn_dh(int i) const { return gdFm ? __ldg(gdFm + i) : 0.; }

    // One barrier per reduction: consecutive reductions alternate between two scratch rows, so a row is
    // rewritten only after every thread has passed the barrier of the reduction in between.
    template <class F> __device__ __forceinline__ double reduce_max(F&& f) {
        double v = f(tid);
        if (v != v) v = -INFINITY;  // skip NaN like `if (x > max)`
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        double* row = redA + flip * kNW;
        flip ^= 1;
        if (lane == 0) row[warp] = v;
        __syncthreads();
        double r = row[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r = fmax(r, __shfl_xor_sync(0xffffffffu, r, o));
        return r;
    }
    template <class F> __device__ __forceinline__ int reduce_max_int(F&& f) {
        int v = f(tid);
        v = __reduce_max_sync(0xffffffffu, v);
        int* row = reinterpret_cast<int*>(redA + flip * kNW);
        flip ^= 1;
        if (lane == 0) row[warp] = v;
        __syncthreads();
        return __reduce_max_sync(0xffffffffu, row[lane]);
    }
    template <class F> __device__ __forceinline__ int reduce_min_int(F&& f) {
        int v = f(tid);
        v = __reduce_min_sync(0xffffffffu, v);
        int* row = reinterpret_cast<int*>(redA + flip * kNW);
        flip ^= 1;
        if (lane == 0) row[warp] = v;
        __syncthreads();
        return __reduce_min_sync(0xffffffffu, row[lane]);
    }
    template <class F> __device__ __forceinline__ void reduce_sums(int nq, F&& f) {
        for (int q = 0; q < nq; ++q) {
            double v = f(tid, q);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) redQ[warp * FB200_MAXQ + q] = v;
        }
        __syncthreads();
        if (tid < nq) {
            double s = 0.;
#pragma unroll 8
            for (int w = 0; w < kNW; ++w) s += redQ[w * FB200_MAXQ + tid];
            sums[tid] = s;
        }
        __syncthreads();
    }
};

// Loads one model's constants, grid and state into shared memory and builds the engine inputs.
// Returns the CTA-uniform ModelState.  All threads must call it.
__device__ __forceinline__ void load_model(const EnsembleDev& E, int model, fb200_model_dev_t* Msh, const SharedArrays& S, GpuEx& ex,
                                           const double* F_src) {
    const int tid = threadIdx.x;
    {
        const double* src = reinterpret_cast<const double*>(E.models + model);
        double* dst = reinterpret_cast<double*>(Msh);
        for (int k = tid; k < int(sizeof(fb200_model_dev_t) / sizeof(double)); k += kNT) dst[k] = src[k];
    }
    const int Nx = E.Nx;
    const size_t off = size_t(model) * Nx;
    if (tid < Nx) {
        S.h[tid] = E.h[off + tid];
        S.F[tid] = F_src[tid];
    } else {
        S.h[tid] = 0.; S.F[tid] = 0.; S.R[tid] = 0.;
    }
    if (tid < E.cfg.n_lambdas) ex.lampref[tid] = FB_DOUBLE_H_C2 / p5(E.cfg.lambdas[tid]);
    ex.gA = E.wA ? E.wA + off : nullptr;
    ex.gB = E.wB ? E.wB + off : nullptr;
    ex.gCs = E.wCs ? E.wCs + off : nullptr;
    ex.gFm = E.Fmagn ? E.Fmagn + off : nullptr;
    ex.gdFm = E.dFmagn_dh ? E.dFmagn_dh + off : nullptr;
    __syncthreads();
}

}  // namespace fb200
