// Read-out kernels: radial fields, spectral flux densities and Mdot_wind of a recorded or the current state.
// They restate the array getters of FreddiState (cpp/src/freddi_state.cpp:184-312), flux_region<Region>(lambda)
// (cpp/include/freddi_state.hpp:377-407) and Mdot_wind (cpp/src/freddi_state.cpp:364-383) on top of the same
// per-point device functions the evolve kernel uses.  One CTA per model.
#include "gpu_exec.cuh"

namespace fb200 {
namespace {

struct Loaded {
    int first, last;
    bool valid;  // false: the requested row was never reached by this model (collapsed earlier)
};

// Stages model `model` at `row` (row < 0: current state).  Returns first/last of that state.
__device__ __forceinline__ Loaded stage(const EnsembleDev& E, int model, long long row, GpuArrays& S, GpuEx& ex) {
    const int Nx = E.Nx;
    const double* F_src = E.F + size_t(model) * Nx;
    Loaded L;
    const ModelState st = E.state[model];
    L.first = st.first;
    L.last = st.last;
    if (row >= 0 && E.Fsnap) {
        F_src = E.Fsnap + (size_t(model) * E.cfg.n_rows + size_t(row)) * Nx;
        const int* b = E.bounds + (size_t(model) * E.cfg.n_rows + size_t(row)) * 2;
        L.first = b[0];
        L.last = b[1];
    }
    L.valid = (L.first >= 0 && L.last > L.first && L.last < Nx);
    if (!L.valid) { L.first = 0; L.last = Nx - 1; F_src = E.F + size_t(model) * Nx; }
    S.h = E.h + size_t(model) * Nx;
    load_model(E, model, S, ex, F_src);
    return L;
}

// The update() whose value ring i of a dynamic wind holds at `row` (cpp/src/freddi_state.cpp:457-654: update() rewrites the
// rings of [first, last] only, so a ring outside keeps what the last update that covered it wrote; `first` only grows
// and `last` only shrinks, so that is the latest record at or before `row` whose range holds i; record 0 is the
// constructor's).  False: the ring was never written (BasicWind zero-initialises A, B, C).
__device__ __forceinline__ bool wind_record_of(const EnsembleDev& E, int model, int row, int i, WindRecord* out) {
    const WindRecord* rec = E.wind_rec + size_t(model) * E.cfg.n_rows;
    for (int s = row; s >= 0; --s) {
        const WindRecord r = rec[s];
        if (i >= r.first && i <= r.last) { *out = r; return true; }
    }
    return false;
}

// Work items are (model, row) pairs: rows row0 .. row0 + n_sel - 1 of every model (n_sel = 1, row0 < 0: the current
// state).  out[model][n_sel][Nx]; pad_nan: NaN outside [first, last] like EvolutionResult (python/freddi/evolution_result.py:55-71).
__global__ void __launch_bounds__(kNT, kCtasPerSm) fb200_field_kernel(const EnsembleDev E, int field, long long row0, int n_sel, int pad_nan,
                                                             double* out) {
    GpuArrays S = carve_arrays(E.cta_scratch + size_t(blockIdx.x) * SmemLayout::kScratchDoublesPerCta, E.h);
    GpuEx ex;
    ex.init();
    ex.set_nx(E.Nx);
    const OutCfg& cfg = stage_cfg(E);
    const int Nx = E.Nx;
    for (long long item = blockIdx.x; item < (long long)E.n_models * n_sel; item += gridDim.x) {
        const int model = int(item / n_sel);
        const long long row = row0 + (item - (long long)model * n_sel);
        __syncthreads();
        const Loaded L = stage(E, model, row, S, ex);
        const fb200_model_dev_t& M = *smem_model();
        ModelState st{};
        st.first = L.first; st.last = L.last;
        DiscEngine<GpuEx> eng(ex, M, cfg, S, st);
        eng.init_points();
        const int first = L.first, last = L.last;
        const StateScalars sc = state_scalars(ex, M, S, first, true);
        for (int i = threadIdx.x; i < Nx; i += kNT) {
            double v = 0.;
            const double h = S.h[i], R = S.R[i], F = S.F[i];
            const bool hot = (i >= first && i <= last);
            PointStructure q{};
            if (hot) q = hot_structure(M, sc, h, R, S.hn[i], S.hm175[i], S.hotR[i], F);
            // cold zone (freddi_state.cpp:256-258,271-273): Height = h2rcold R, Kirr from the cold parameters
            const double Hc = M.h2rcold * R;
            const double Kc = M.Cirr_cold * ((M.irrindex_cold == 0.) ? 1.0 : pow(Hc / (R * 0.05), M.irrindex_cold));
            const double H = hot ? q.Height : Hc;
            const double K = hot ? q.Kirr : Kc;
            double Qx = 0.;
            if (i >= first) {
                const double mu = H / R;
                if (M.is_ns) Qx = K * (sc.Lbol_disk * angular_dist(M.angdist_disk, mu) + sc.Lbol_ns * angular_dist(M.angdist_ns, mu)) / (4. * FB_PI * p2(R));
                else Qx = K * sc.Lbol_disk * angular_dist(M.angdist_disk, mu) / (4. * FB_PI * p2(R));
            }
            switch (field) {
                case FB200_FIELD_H: v = h; break;
                case FB200_FIELD_R: v = R; break;
                case FB200_FIELD_F: v = F; break;
                case FB200_FIELD_W: v = hot ? q.W : 0.; break;
                case FB200_FIELD_SIGMA: v = hot ? q.Sigma : 0.; break;
                case FB200_FIELD_TPH_VIS: v = hot ? q.Tvis : 0.; break;
                case FB200_FIELD_TPH: v = (i >= first) ? pow025(p4(hot ? q.Tvis : 0.) + Qx / FB_SIGMA_SB) : 0.; break;
                case FB200_FIELD_TIRR: v = (i >= first) ? pow025(Qx / FB_SIGMA_SB) : 0.; break;
                case FB200_FIELD_KIRR: v = (i >= first) ? K : 0.; break;
                case FB200_FIELD_HEIGHT: v = (i >= first) ? H : 0.; break;
                case FB200_FIELD_QX: v = Qx; break;
                case FB200_FIELD_TPH_X: {
                    v = 0.;
                    if (hot) {
                        const double KM = 3. * fabs(sc.Mdot_plain) * p6(FB_C_LIGHT) / (8. * FB_PI * p2(M.GM));
                        double x = S.c1[i] * S.F[first];
                        const double g = KM * S.Gb[i];
                        x += (g < 0.) ? FB_NAN_INVALID : g;
                        v = M.colourfactor * pow025(x / FB_SIGMA_SB);
                    }
                    break;
                }
                case FB200_FIELD_WINDA: v = ex.windA(i); break;
                case FB200_FIELD_WINDB:
                case FB200_FIELD_WINDC: {
                    double B = ex.windB(i), C = ex.windC_static(i);
                    if (M.wind_dynamic && E.wind_rec) {
                        // dynamic winds: the reference's getters return what the last update() left — computed inside step()
                        // BEFORE the solve, from the previous state's Mdot_in, and only for the rings of that update's [first, last]
                        WindRecord rec;
                        const int at = (row >= 0) ? int(row) : E.state[model].i_t;
                        if (wind_record_of(E, model, at, i, &rec)) {
                            const WindScalars ws = wind_scalars(M, rec.Mdot, S.h[rec.first], S.h[rec.last], S.h[Nx - 1]);
                            double Bd = 0., Cd = 0.;
                            wind_dynamic_point(M, ws, h, R, &Bd, &Cd);
                            if (M.windtype == FB200_WIND_JANIUK2015) B = Bd; else C = Cd + C;
                        }
                    }
                    v = (field == FB200_FIELD_WINDB) ? B : C;
                    break;
                }
                case FB200_FIELD_FMAGN: v = ex.Fmagn(i); break;
                case FB200_FIELD_DFMAGN_DH: v = ex.dFmagn_dh(i); break;
                case FB200_FIELD_D2FMAGN_DH2: v = E.d2Fmagn_dh2 ? E.d2Fmagn_dh2[size_t(model) * Nx + i] : 0.; break;
                default: v = NAN;
            }
            if (pad_nan && !(i >= first && i <= last)) v = NAN;
            out[size_t(item) * Nx + i] = L.valid ? v : NAN;
        }
    }
}

// out[model][k] = flux_region<region>(lambdas[k]); region 0 hot (Tph over [first,last]), 1 cold (Tirr over [last+1,Nx-1])
// region 2: the irradiated companion star seen at orbital phase phases[row - row0] (FreddiState::flux_star(lambda, phase),
// cpp/src/freddi_state.cpp:355-357).  Work items (model, row) as in fb200_field_kernel; out[model][n_sel][n_lambdas].
__global__ void __launch_bounds__(kNT, kCtasPerSm) fb200_flux_kernel(const EnsembleDev E, int region, long long row0, int n_sel,
                                                            const double* lambdas, int n_lambdas, const double* phases, double* out) {
    GpuArrays S = carve_arrays(E.cta_scratch + size_t(blockIdx.x) * SmemLayout::kScratchDoublesPerCta, E.h);
    GpuEx ex;
    ex.init();
    ex.set_nx(E.Nx);
    const OutCfg& cfg = stage_cfg(E);
    const int Nx = E.Nx;
    for (long long item = blockIdx.x; item < (long long)E.n_models * n_sel; item += gridDim.x) {
        const int model = int(item / n_sel);
        const long long row = row0 + (item - (long long)model * n_sel);
        __syncthreads();
        const Loaded L = stage(E, model, row, S, ex);
        const fb200_model_dev_t& M = *smem_model();
        ModelState st{};
        st.first = L.first; st.last = L.last;
        DiscEngine<GpuEx> eng(ex, M, cfg, S, st);
        eng.init_points();
        const int first = L.first, last = L.last;
        const StateScalars sc = state_scalars(ex, M, S, first, true);
        if (region == 2) {
            // Height2R of the irradiation sources (freddi_state.cpp:663-671), then Teff of every triangle, then one
            // directional sum per wavelength.  The sources exist from the first step on (freddi_evolution.cpp:28).
            const double h2r = ex.reduce_max([&](int i, int) -> double {
                if (i < first || i >= Nx) return 0.;
                const double R = S.R[i];
                if (i <= last) return hot_structure(M, sc, S.h[i], R, S.hn[i], S.hm175[i], S.hotR[i], S.F[i]).Height / R;
                return (M.h2rcold * R) / R;
            });
            const bool with_sources = (row >= 0) ? (row > 0) : (E.state[model].i_t > 0);
            __syncthreads();
            eng.star_teff(sc, (h2r > 0.) ? h2r : 0., with_sources);
            const double phase = phases[item - (long long)model * n_sel];
            for (int k = 0; k < n_lambdas; ++k) {
                const double lam = lambdas[k];
                eng.star_dir_sums(lam, FB_DOUBLE_H_C2 / p5(lam), -1, phase);
                if (threadIdx.x == 0)
                    out[size_t(item) * n_lambdas + k] = L.valid ? ex.sum(0) * (4. * FB_PI) / (4. * FB_PI * p2(M.distance)) : NAN;
                __syncthreads();
            }
            continue;
        }
        // temperature of the region at every point
        ex.points([&](int i, int) {
            double T = 0.;
            if (i < Nx) {
                const double R = S.R[i];
                if (region == 0) {
                    if (i >= first && i <= last) T = hot_structure(M, sc, S.h[i], R, S.hn[i], S.hm175[i], S.hotR[i], S.F[i]).Tph;
                } else if (i > last) {
                    const double Hc = M.h2rcold * R;
                    const double Kc = M.Cirr_cold * ((M.irrindex_cold == 0.) ? 1.0 : pow(Hc / (R * 0.05), M.irrindex_cold));
                    const double mu = Hc / R;
                    double Qx;
                    if (M.is_ns) Qx = Kc * (sc.Lbol_disk * angular_dist(M.angdist_disk, mu) + sc.Lbol_ns * angular_dist(M.angdist_ns, mu)) / (4. * FB_PI * p2(R));
                    else Qx = Kc * sc.Lbol_disk * angular_dist(M.angdist_disk, mu) / (4. * FB_PI * p2(R));
                    T = pow025(Qx / FB_SIGMA_SB);
                }
            }
            S.Tph[i] = T;
        });
        ex.sync();
        const int rf = (region == 0) ? first : last + 1;
        const int rl = (region == 0) ? last : Nx - 1;
        for (int k0 = 0; k0 < n_lambdas; k0 += FB200_MAXQ) {
            const int nq = min(FB200_MAXQ, n_lambdas - k0);
            ex.reduce_sums(nq, [&](int i, int, int q) -> double {
                if (rf >= rl || i < rf || i > rl) return 0.;
                const double w = trapz_weight(S.R, i, rf, rl);
                const double lam = lambdas[k0 + q];
                const double y = planck_lambda_pref(S.Tph[i], lam, FB_DOUBLE_H_C2 / p5(lam));
                return 2 * FB_PI * S.R[i] * y * w;
            });
            if (threadIdx.x < nq) {
                const double lam = lambdas[k0 + threadIdx.x];
                out[size_t(item) * n_lambdas + k0 + threadIdx.x] =
                    L.valid ? 0.5 * ex.sum(threadIdx.x) * p2(lam) / FB_C_LIGHT * M.cosiOverD2 : NAN;
            }
            __syncthreads();
        }
    }
}

// Mdot_wind (cpp/src/freddi_state.cpp:364-383): - trapz over h of (A dF/dh + B F + C) on [first, last]
__global__ void __launch_bounds__(kNT, kCtasPerSm) fb200_mdot_wind_kernel(const EnsembleDev E, long long row, double* out) {
    GpuArrays S = carve_arrays(E.cta_scratch + size_t(blockIdx.x) * SmemLayout::kScratchDoublesPerCta, E.h);
    GpuEx ex;
    ex.init();
    ex.set_nx(E.Nx);
    const OutCfg& cfg = stage_cfg(E);
    const int Nx = E.Nx;
    for (int model = blockIdx.x; model < E.n_models; model += gridDim.x) {
        __syncthreads();
        const Loaded L = stage(E, model, row, S, ex);
        const fb200_model_dev_t& M = *smem_model();
        ModelState st{};
        st.first = L.first; st.last = L.last;
        DiscEngine<GpuEx> eng(ex, M, cfg, S, st);
        eng.init_points();
        const int first = L.first, last = L.last;
        const StateScalars sc = state_scalars(ex, M, S, first, false);
        // the wind arrays behind Mdot_wind() are those of the last update(): the record of this row covers [first, last]
        WindScalars ws{0, 0, 0, 0, 0};
        bool dyn = false;
        if (M.wind_dynamic && E.wind_rec) {
            const WindRecord rec = E.wind_rec[size_t(model) * E.cfg.n_rows + ((row >= 0) ? int(row) : E.state[model].i_t)];
            if (rec.first <= rec.last) {
                dyn = true;
                ws = wind_scalars(M, rec.Mdot, S.h[rec.first], S.h[rec.last], S.h[Nx - 1]);
            }
        }
        ex.reduce_sums(1, [&](int i, int, int) -> double {
            if (first >= last || i < first || i > last) return 0.;
            const double* h = S.h;
            const auto& F = S.F;
            double dFdh;
            if (i == first) {
                dFdh = (F[i + 1] - F[i]) / (h[i + 1] - h[i]);
            } else if (i == last) {
                dFdh = (F[i] - F[i - 1]) / (h[i] - h[i - 1]);
            } else {
                const double delta_0 = h[i] - h[i - 1];
                const double delta_1 = h[i + 1] - h[i];
                dFdh = (F[i + 1] * delta_0 * delta_0 / (delta_0 + delta_1) + F[i] * (delta_1 - delta_0) -
                        F[i - 1] * delta_1 * delta_1 / (delta_0 + delta_1)) /
                       (delta_0 * delta_1);
            }
            double A = ex.windA(i), B = ex.windB(i), C = ex.windC_static(i);
            if (dyn) {
                double Bd = 0., Cd = 0.;
                wind_dynamic_point(M, ws, h[i], S.R[i], &Bd, &Cd);
                if (M.windtype == FB200_WIND_JANIUK2015) B = Bd; else C = Cd + C;
            }
            const double y = -(A * dFdh + B * F[i] + C);
            return y * trapz_weight(h, i, first, last);
        });
        if (threadIdx.x == 0) out[model] = L.valid ? 0.5 * ex.sum(0) : NAN;
    }
}

// fb200_selftest_trapz: the CTA-wide trapezoid of the kernels (policy_trapz) on caller data, x in S.R, y in S.F
__global__ void __launch_bounds__(kNT, kCtasPerSm) fb200_trapz_kernel(const double* x, const double* y, int n, int first, int last, double* out,
                                                             double* gscratch) {
    GpuArrays S = carve_arrays(gscratch, x);
    GpuEx ex;
    ex.init();
    ex.set_nx(n);
    for (int i = threadIdx.x; i < FB200_NX_MAX; i += kNT) {
        S.R[i] = (i < n) ? x[i] : 0.;
        S.F[i] = (i < n) ? y[i] : 0.;
    }
    __syncthreads();
    const double v = policy_trapz(ex, S.R, S.F, first, last);
    if (threadIdx.x == 0) out[0] = v;
}

// 16 independent DFMA chains per thread, 4096 iterations: 2 * 16 * 4096 flop per thread
__global__ void __launch_bounds__(256) fb200_dfma_peak_kernel(double* out, double a, double b) {
    double x[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = a + k + threadIdx.x;
    for (int it = 0; it < 4096; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = fma(x[k], a, b);
    }
    double s = 0.;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += x[k];
    if (s == 123.456) out[0] = s;  // keeps the chains alive
}

cudaError_t prep(const void* fn) { return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SmemLayout::bytes())); }

}  // namespace

cudaError_t measure_fp64_peak(double* tflops) {
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double* d = nullptr;
    e = cudaMalloc(&d, 64);
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int grid = sms * 32, block = 256;
    double best_ms = 1e30;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        fb200_dfma_peak_kernel<<<grid, block>>>(d, 0.999999, 1e-7);
        cudaEventRecord(e1);
        e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best_ms) best_ms = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    if (e != cudaSuccess) return e;
    const double flop = 2.0 * 16 * 4096 * double(grid) * block;
    *tflops = flop / (best_ms * 1e-3) / 1e12;
    return cudaGetLastError();
}

#ifndef FB_VARIANT
cudaError_t selftest_trapz(const double* x_dev, const double* y_dev, int n, int first, int last, double* out_dev, double* scratch_dev,
                           cudaStream_t stream) {
    cudaError_t e = prep(reinterpret_cast<const void*>(fb200_trapz_kernel));
    if (e != cudaSuccess) return e;
    fb200_trapz_kernel<<<1, kNT, SmemLayout::bytes(), stream>>>(x_dev, y_dev, n, first, last, out_dev, scratch_dev);
    return cudaGetLastError();
}
#endif

cudaError_t launch_field(const EnsembleDev& E, int field, long long row0, int n_sel, int pad_nan, double* out, cudaStream_t stream) {
    cudaError_t e = prep(reinterpret_cast<const void*>(fb200_field_kernel));
    if (e != cudaSuccess) return e;
    const long long items = (long long)E.n_models * n_sel;
    const int grid = items < E.max_ctas ? int(items) : E.max_ctas;
    fb200_field_kernel<<<grid, kNT, SmemLayout::bytes(), stream>>>(E, field, row0, n_sel, pad_nan, out);
    return cudaGetLastError();
}

cudaError_t launch_flux(const EnsembleDev& E, int region, long long row0, int n_sel, const double* lambdas_dev, int n_lambdas,
                        const double* phases_dev, double* out, cudaStream_t stream) {
    cudaError_t e = prep(reinterpret_cast<const void*>(fb200_flux_kernel));
    if (e != cudaSuccess) return e;
    const long long items = (long long)E.n_models * n_sel;
    const int grid = items < E.max_ctas ? int(items) : E.max_ctas;
    fb200_flux_kernel<<<grid, kNT, SmemLayout::bytes(), stream>>>(E, region, row0, n_sel, lambdas_dev, n_lambdas, phases_dev, out);
    return cudaGetLastError();
}

cudaError_t launch_mdot_wind(const EnsembleDev& E, long long row, double* out, cudaStream_t stream) {
    cudaError_t e = prep(reinterpret_cast<const void*>(fb200_mdot_wind_kernel));
    if (e != cudaSuccess) return e;
    const int grid = E.n_models < E.max_ctas ? E.n_models : E.max_ctas;
    fb200_mdot_wind_kernel<<<grid, kNT, SmemLayout::bytes(), stream>>>(E, row, out);
    return cudaGetLastError();
}

}  // namespace fb200

#ifndef FB_VARIANT
#define FB_VARIANT main
#endif
#define FB_CAT3_(a, b, c) a##b##c
#define FB_CAT3(a, b, c) FB_CAT3_(a, b, c)
#define FB_VNAME(name) FB_CAT3(fb200_, FB_VARIANT, name)
extern "C" {  // see evolve.cu
cudaError_t FB_VNAME(_launch_field)(const void* E, int field, long long row0, int n_sel, int pad_nan, double* out, cudaStream_t stream) {
    return fb200::launch_field(*static_cast<const fb200::EnsembleDev*>(E), field, row0, n_sel, pad_nan, out, stream);
}
cudaError_t FB_VNAME(_launch_flux)(const void* E, int region, long long row0, int n_sel, const double* lambdas_dev, int n_lambdas,
                                   const double* phases_dev, double* out, cudaStream_t stream) {
    return fb200::launch_flux(*static_cast<const fb200::EnsembleDev*>(E), region, row0, n_sel, lambdas_dev, n_lambdas, phases_dev, out, stream);
}
cudaError_t FB_VNAME(_launch_mdot_wind)(const void* E, long long row, double* out, cudaStream_t stream) {
    return fb200::launch_mdot_wind(*static_cast<const fb200::EnsembleDev*>(E), row, out, stream);
}
}
