// extern "C" layer of include/freddi_b200.h: owns device memory, resolves arguments on the host, launches
// the kernels.  No CPU implementation of the path lives here: without a CUDA device every compute call fails.
#include <atomic>
#include <memory>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "evolve.cuh"
#include "resolve.hpp"
#include "star.hpp"

using namespace fb200;

// the large-grid build of the kernels (evolve_big.cu, fields_big.cu)
extern "C" {
size_t fb200_big_evolve_smem_bytes();
size_t fb200_big_evolve_scratch_doubles_per_cta();
int fb200_big_evolve_ctas_per_sm();
int fb200_big_evolve_resident_ctas_per_sm();
cudaError_t fb200_big_launch_evolve(const void* E, int n_ctas, cudaStream_t stream);
cudaError_t fb200_big_launch_field(const void* E, int field, long long row0, int n_sel, int pad_nan, double* out, cudaStream_t stream);
cudaError_t fb200_big_launch_flux(const void* E, int region, long long row0, int n_sel, const double* lambdas_dev, int n_lambdas,
                                  const double* phases_dev, double* out, cudaStream_t stream);
cudaError_t fb200_big_launch_mdot_wind(const void* E, long long row, double* out, cudaStream_t stream);
}

namespace {
thread_local std::string g_err;
int set_err(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? FB200_ERR_NO_DEVICE : FB200_ERR_CUDA;
}
#define FB_CUDA(call)                                         \
    do {                                                      \
        cudaError_t _e = (call);                              \
        if (_e != cudaSuccess) return cuda_fail(_e, #call);   \
    } while (0)
}  // namespace

struct fb200_ensemble {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    size_t n = 0;
    int Nx = 0, n_rows = 0, n_cols = 0, sm_count = 0, max_Nt = 0, chunk_steps = 25;
    bool is_ns = false;
    bool big = false;  // Nx > 1024: the large-grid build of the kernels
    fb200_config_t cfg{};
    std::vector<fb200_model_info_t> info;
    EnsembleDev dev{};
    std::vector<void*> allocs;
    int last_launches = 0;
    bool timed = false;
    double* scratch = nullptr;  // [n * Nx] read-out staging
    size_t scratch_len = 0;
    // initial conditions kept on the host for fb200_ensemble_reset (pinned staging is allocated on first use)
    std::unique_ptr<double[]> host_F0;  // [n * Nx]
    std::vector<ModelState> host_state0;
    double* pinned_F0 = nullptr;
    ModelState* pinned_state0 = nullptr;
};

namespace {

template <class T>
int dev_alloc(fb200_ensemble* e, T** p, size_t count) {
    void* q = nullptr;
    FB_CUDA(cudaMalloc(&q, count * sizeof(T) + 16));
    e->allocs.push_back(q);
    *p = static_cast<T*>(q);
    return FB200_OK;
}

int select_device(const fb200_ensemble* e) {
    FB_CUDA(cudaSetDevice(e->device));
    return FB200_OK;
}

int readout_scratch_init(fb200_ensemble* e, size_t len, double** p) {
    FB_CUDA(cudaMalloc(reinterpret_cast<void**>(&e->scratch), len * sizeof(double)));
    e->scratch_len = len;
    *p = e->scratch;
    return FB200_OK;
}

}  // namespace

extern "C" {

int fb200_abi_version(void) { return FB200_ABI_VERSION; }
const char* fb200_last_error(void) { return g_err.c_str(); }

int fb200_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cuda_fail(e, "cudaGetDeviceCount");
        return (e == cudaErrorNoDevice) ? 0 : -1;
    }
    return n;
}

int fb200_measure_fp64_peak(int device, double* tflops) {
    if (!tflops) return set_err(FB200_ERR_INVALID_ARGUMENT, "null argument");
    FB_CUDA(cudaSetDevice(device));
    double tf = 0.;
    cudaError_t ce = fb200::measure_fp64_peak(&tf);
    if (ce != cudaSuccess) return cuda_fail(ce, "fp64 peak kernel");
    *tflops = tf;
    return FB200_OK;
}

int fb200_selftest_trapz(int device, const double* x, const double* y, size_t n, size_t first, size_t last, double* out) {
    if (!x || !y || !out) return set_err(FB200_ERR_INVALID_ARGUMENT, "null argument");
    if (n < 2 || n > 1024 || last >= n) return set_err(FB200_ERR_INVALID_ARGUMENT, "need 2 <= n <= 1024 points and last < n");
    FB_CUDA(cudaSetDevice(device));
    double* d = nullptr;
    const size_t slot = evolve_scratch_doubles_per_cta();
    FB_CUDA(cudaMalloc(reinterpret_cast<void**>(&d), (2 * n + 1 + slot) * sizeof(double)));
    cudaError_t ce = cudaMemcpy(d, x, n * sizeof(double), cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemcpy(d + n, y, n * sizeof(double), cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = selftest_trapz(d, d + n, int(n), int(first), int(last), d + 2 * n, d + 2 * n + 1, nullptr);
    if (ce == cudaSuccess) ce = cudaMemcpy(out, d + 2 * n, sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (ce != cudaSuccess) return cuda_fail(ce, "fb200_trapz_kernel");
    return FB200_OK;
}

// defaults of cpp/pywrap/pywrap_freddi_evolution.cpp:33-83,203-224
void fb200_args_default(fb200_args_t* a, int is_ns) {
    std::memset(a, 0, sizeof(*a));
    a->is_ns = is_ns ? 1 : 0;
    const double nan = NAN;
    a->alpha = nan; a->alphacold = nan; a->Mx = nan; a->kerr = 0.0; a->period = nan; a->Mopt = nan;
    a->rochelobefill = 1.0; a->Topt = 0.0; a->rin = nan; a->rout = nan; a->risco = nan;
    a->opacity = FB200_OPACITY_KRAMERS; a->boundcond = FB200_BOUNDCOND_TEFF; a->initialcond = FB200_IC_POWERF;
    a->windtype = FB200_WIND_NO;
    a->Mdotout = 0.0; a->Thot = 0.0; a->Qirr2Qvishot = 0.0;
    a->F0 = nan; a->Mdisk0 = nan; a->Mdot0 = nan; a->powerorder = nan; a->gaussmu = nan; a->gausssigma = nan;
    a->Cirr = 0.0; a->irrindex = 0.0; a->Cirrcold = 0.0; a->irrindexcold = 0.0; a->h2rcold = 0.0;
    a->angulardistdisk = FB200_ANGDIST_PLANE; a->angulardistns = FB200_ANGDIST_ISOTROPIC;
    a->colourfactor = 1.7;
    a->emin = 1.0 * 1000. * FB_ELECTRON_VOLT / FB_H_PLANCK;   // kevToHertz(1.)
    a->emax = 12.0 * 1000. * FB_ELECTRON_VOLT / FB_H_PLANCK;  // kevToHertz(12.)
    a->staralbedo = 0.0; a->inclination = 0.0; a->ephemerist0 = 0.0; a->distance = nan;
    a->inittime = 0.0; a->time = nan; a->tau = nan; a->eps = 1e-6;
    a->Nx = 1000; a->gridscale = FB200_GRID_LOG; a->starlod = 3; a->nsprop = FB200_NSPROP_DUMMY;
    a->freqx = nan; a->Rx = nan; a->Bx = nan; a->hotspotarea = 1.0; a->epsilonAlfven = 1.0; a->inversebeta = 0.0; a->Rdead = 0.0;
    a->fptype = FB200_FP_NO_OUTFLOW; a->kappattype = FB200_KAPPAT_CONST; a->nsgravredshift = FB200_REDSHIFT_OFF;
    a->kappatparams[0] = 1.0 / 3.0; a->kappatparams[1] = 1.0 / 3.0;
}

void fb200_config_default(fb200_config_t* c) {
    std::memset(c, 0, sizeof(*c));
    c->compute_lx = 1;
}

int fb200_args_resolve(const fb200_args_t* args, fb200_model_info_t* info, double* h, double* F) {
    if (!args) return set_err(FB200_ERR_INVALID_ARGUMENT, "null argument");
    Resolved r;
    std::string msg;
    const int rc = resolve(*args, r, msg);
    if (rc != FB200_OK) return set_err(rc, msg);
    if (info) *info = r.info;
    if (h) std::memcpy(h, r.h.data(), sizeof(double) * r.h.size());
    if (F) std::memcpy(F, r.F.data(), sizeof(double) * r.F.size());
    return FB200_OK;
}

int fb200_ensemble_create(const fb200_args_t* args, size_t n_models, const fb200_config_t* cfg_in, fb200_ensemble_t** out,
                          size_t* bad_model) {
    if (bad_model) *bad_model = 0;
    if (!args || !out || n_models == 0) return set_err(FB200_ERR_INVALID_ARGUMENT, "null argument or empty ensemble");
    const bool dbg_time = getenv("FB200_DEBUG") != nullptr;  // FB200_DEBUG: where the time of a create goes, on stderr
    auto dbg_t0 = std::chrono::steady_clock::now();
    auto dbg_lap = [&](const char* what) {
        if (!dbg_time) return;
        const auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[fb200] create: %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t - dbg_t0).count());
        dbg_t0 = t;
    };
    fb200_config_t cfg;
    if (cfg_in) cfg = *cfg_in; else fb200_config_default(&cfg);
    if (cfg.n_lambdas < 0 || cfg.n_lambdas > FB200_MAX_LAMBDAS) return set_err(FB200_ERR_INVALID_ARGUMENT, "n_lambdas out of range");
    if (cfg.n_passbands < 0 || cfg.n_passbands > FB200_MAX_PASSBANDS) return set_err(FB200_ERR_INVALID_ARGUMENT, "n_passbands out of range");
    std::vector<PassbandTable> passbands(cfg.n_passbands);
    for (int p = 0; p < cfg.n_passbands; ++p) {
        std::string msg;
        const int rc = prepare_passband(cfg.passband_lambdas[p], cfg.passband_transmissions[p], cfg.passband_len[p], passbands[p], msg);
        if (rc != FB200_OK) return set_err(rc, "passband " + std::to_string(p) + ": " + msg);
    }

    // ---- resolve every model on the host ----
    std::vector<Resolved> res(n_models);
    std::vector<int> codes(n_models, FB200_OK);
    std::vector<std::string> errs(n_models);
    // body(i) for every model, on the host threads of the configuration (models are handed out one by one: their cost varies)
    auto for_each_model = [&](auto&& body) {
        unsigned nt = cfg.host_threads > 0 ? unsigned(cfg.host_threads) : std::thread::hardware_concurrency();
        if (nt == 0) nt = 1;
        if (nt > n_models) nt = unsigned(n_models);
        std::atomic<size_t> next{0};
        auto work = [&] {
            for (size_t i; (i = next.fetch_add(1)) < n_models;) body(i);
        };
        if (nt <= 1) {
            work();
        } else {
            std::vector<std::thread> pool;
            for (unsigned t = 0; t < nt; ++t) pool.emplace_back(work);
            for (auto& th : pool) th.join();
        }
    };
    for_each_model([&](size_t i) { codes[i] = resolve(args[i], res[i], errs[i]); });
    dbg_lap("host resolve");
    for (size_t i = 0; i < n_models; ++i) {
        if (codes[i] != FB200_OK) {
            if (bad_model) *bad_model = i;
            return set_err(codes[i], errs[i]);
        }
    }
    const int Nx = res[0].dev.Nx;
    const bool is_ns = res[0].dev.is_ns != 0;
    int max_Nt = 0;
    bool anyA = false, anyB = false, anyC = false;
    for (size_t i = 0; i < n_models; ++i) {
        if (res[i].dev.Nx != Nx || (res[i].dev.is_ns != 0) != is_ns) {
            if (bad_model) *bad_model = i;
            return set_err(FB200_ERR_INVALID_ARGUMENT, "all models of one ensemble must share Nx and the model class (BH / NS)");
        }
        max_Nt = std::max(max_Nt, res[i].dev.Nt);
        anyA |= !res[i].wA.empty();
        anyB |= !res[i].wB.empty();
        anyC |= !res[i].wC.empty();
        if (is_ns && res[i].dev.inverse_beta != 0.) anyC = true;
    }

    // ---- device ----
    int ndev = 0;
    {
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0) {
            if (e != cudaSuccess) cuda_fail(e, "cudaGetDeviceCount"); else g_err = "no CUDA device";
            return set_err(FB200_ERR_NO_DEVICE, "no CUDA device: " + g_err + " (there is no CPU fallback)");
        }
    }
    if (cfg.device < 0 || cfg.device >= ndev) return set_err(FB200_ERR_INVALID_ARGUMENT, "device ordinal out of range");

    auto* e = new fb200_ensemble();
    auto guard_fail = [&](int code) { fb200_ensemble_destroy(e); return code; };
    e->device = cfg.device;
    e->cfg = cfg;
    for (int p = 0; p < FB200_MAX_PASSBANDS; ++p) { e->cfg.passband_lambdas[p] = nullptr; e->cfg.passband_transmissions[p] = nullptr; }  // caller-owned, read above
    e->n = n_models;
    e->Nx = Nx;
    e->is_ns = is_ns;
    e->n_rows = max_Nt + 1;
    e->max_Nt = max_Nt;
    if (const char* c = getenv("FB200_CHUNK_STEPS")) e->chunk_steps = atoi(c) > 0 ? atoi(c) : 0x3fffffff;
    e->n_cols = FB200_N_BH_COLS + (cfg.n_lambdas + cfg.n_passbands) * ((cfg.cold_disk ? 2 : 1) + (cfg.star_flux == 1 ? 3 : 0)) +
                (is_ns ? FB200_N_NS_COLS : 0);
    e->info.resize(n_models);
    for (size_t i = 0; i < n_models; ++i) e->info[i] = res[i].info;
    {
        int rc = select_device(e);
        if (rc != FB200_OK) return guard_fail(rc);
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, e->device) != cudaSuccess) return guard_fail(set_err(FB200_ERR_CUDA, "cudaGetDeviceProperties"));
    e->sm_count = prop.multiProcessorCount;
    e->big = Nx > 1024;
    const bool big = e->big;
    if (size_t(prop.sharedMemPerBlockOptin) < (big ? fb200_big_evolve_smem_bytes() : evolve_smem_bytes()))
        return guard_fail(set_err(FB200_ERR_UNSUPPORTED, "device offers too little shared memory per block for the sm_100a kernels"));
#define FB_TRY(x) do { int _rc = (x); if (_rc != FB200_OK) return guard_fail(_rc); } while (0)
#define FB_TRYC(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) return guard_fail(cuda_fail(_e, #call)); } while (0)
    FB_TRYC(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    FB_TRYC(cudaEventCreate(&e->ev0));
    FB_TRYC(cudaEventCreate(&e->ev1));

    EnsembleDev& D = e->dev;
    D.n_models = int(n_models);
    D.Nx = Nx;
    D.cfg.n_lambdas = cfg.n_lambdas;
    D.cfg.cold_disk = cfg.cold_disk;
    D.cfg.compute_lx = cfg.compute_lx;
    D.cfg.record_F = cfg.record_F;
    D.cfg.lx_prune = cfg.exact_lx_walk ? 0 : 1;
    D.cfg.n_cols = e->n_cols;
    D.cfg.n_rows = e->n_rows;
    for (int k = 0; k < FB200_MAX_LAMBDAS; ++k) D.cfg.lambdas[k] = (k < cfg.n_lambdas) ? cfg.lambdas[k] : 0.;

    D.cfg.n_passbands = cfg.n_passbands;
    D.cfg.pb_a = nullptr;
    D.cfg.pb_w = nullptr;
    for (int k = 0; k < FB200_MAX_PASSBANDS + 2; ++k) D.cfg.pb_off[k] = 0;
    for (int k = 0; k < FB200_MAX_PASSBANDS; ++k) D.cfg.pb_tdnu[k] = 1.;
    if (cfg.n_passbands > 0) {  // all tables back to back: a_j, then w_j
        std::vector<double> ha, hw;
        for (int p = 0; p < cfg.n_passbands; ++p) {
            D.cfg.pb_off[p] = int(ha.size());
            ha.insert(ha.end(), passbands[p].a.begin(), passbands[p].a.end());
            hw.insert(hw.end(), passbands[p].w.begin(), passbands[p].w.end());
            D.cfg.pb_tdnu[p] = passbands[p].t_dnu;
        }
        for (int p = cfg.n_passbands; p < FB200_MAX_PASSBANDS + 2; ++p) D.cfg.pb_off[p] = int(ha.size());
        double *d_a = nullptr, *d_w = nullptr;
        FB_TRY(dev_alloc(e, &d_a, ha.size()));
        FB_TRY(dev_alloc(e, &d_w, hw.size()));
        FB_TRYC(cudaMemcpy(d_a, ha.data(), sizeof(double) * ha.size(), cudaMemcpyHostToDevice));
        FB_TRYC(cudaMemcpy(d_w, hw.data(), sizeof(double) * hw.size(), cudaMemcpyHostToDevice));
        D.cfg.pb_a = d_a;
        D.cfg.pb_w = d_w;
    }

    // ---- companion star (--starflux): one triangle table per distinct (semi-axis, mass ratio, fill factor, starlod) ----
    D.cfg.star_flux = (cfg.star_flux == 1) ? 1 : 0;  // 2: geometry for fb200_ensemble_flux_rows(region 2) only, no columns
    D.cfg.star_data = nullptr;
    D.star_T = nullptr;
    D.star_ntri_max = 0;
    if (cfg.star_flux) {
        struct Key { double semiaxis, q, fill; int lod; long long off; int ntri; };
        std::vector<Key> keys;
        std::vector<double> table;
        for (size_t i = 0; i < n_models; ++i) {
            const double q = args[i].Mopt / args[i].Mx;
            const Key* hit = nullptr;
            for (const Key& k : keys)
                if (k.semiaxis == res[i].dev.star_semiaxis && k.q == q && k.fill == args[i].rochelobefill && k.lod == args[i].starlod) { hit = &k; break; }
            if (!hit) {
                StarGeometry g;
                std::string msg;
                const int rc = build_star_geometry(res[i].dev.star_semiaxis, q, args[i].rochelobefill, args[i].starlod, g, msg);
                if (rc != FB200_OK) {
                    if (bad_model) *bad_model = i;
                    return guard_fail(set_err(rc, msg));
                }
                keys.push_back(Key{res[i].dev.star_semiaxis, q, args[i].rochelobefill, args[i].starlod, (long long)table.size(), g.n_tri});
                const std::vector<double> packed = g.packed();
                table.insert(table.end(), packed.begin(), packed.end());
                hit = &keys.back();
            }
            res[i].dev.star_off = hit->off;
            res[i].dev.star_ntri = hit->ntri;
            D.star_ntri_max = std::max(D.star_ntri_max, hit->ntri);
        }
        double* d_star = nullptr;
        FB_TRY(dev_alloc(e, &d_star, table.size()));
        FB_TRYC(cudaMemcpy(d_star, table.data(), sizeof(double) * table.size(), cudaMemcpyHostToDevice));
        D.cfg.star_data = d_star;
    }

    dbg_lap("context, stream, tables");
    const size_t NN = n_models * size_t(Nx);
    fb200_model_dev_t* d_models = nullptr;
    ModelState* d_state = nullptr;
    double *d_h = nullptr, *d_F = nullptr, *d_rows = nullptr;
    FB_TRY(dev_alloc(e, &D.K, NN));
    FB_TRY(dev_alloc(e, &d_models, n_models));
    FB_TRY(dev_alloc(e, &d_state, n_models));
    FB_TRY(dev_alloc(e, &d_h, NN));
    FB_TRY(dev_alloc(e, &d_F, NN));
    const size_t rows_len = n_models * size_t(e->n_rows) * e->n_cols;
    FB_TRY(dev_alloc(e, &d_rows, rows_len));
    FB_TRY(dev_alloc(e, &D.queue, 4));
    {
        // the chunked scheduler lets a CTA wait for another one: never launch more CTAs than can be resident
        int resident = big ? fb200_big_evolve_resident_ctas_per_sm() : evolve_resident_ctas_per_sm();
        if (resident < 1) return guard_fail(set_err(FB200_ERR_UNSUPPORTED, "the evolve kernel does not fit on this device"));
        const int designed = big ? fb200_big_evolve_ctas_per_sm() : evolve_ctas_per_sm();
        if (resident > designed) resident = designed;
        if (big && size_t(e->sm_count) * resident > n_models) {  // 3 MB of scratch per CTA: no more CTAs than models
            resident = int((n_models + e->sm_count - 1) / e->sm_count);
        }
        D.max_ctas = e->sm_count * resident;
    }
    FB_TRY(dev_alloc(e, &D.progress, n_models));
    FB_TRY(dev_alloc(e, &D.failed, n_models));
    FB_TRYC(cudaMemset(D.failed, 0, sizeof(int) * n_models));
    {
        long long ms = cfg.wait_timeout_ms > 0 ? cfg.wait_timeout_ms : 10000;
        if (const char* c = getenv("FB200_WAIT_TIMEOUT_US")) D.wait_timeout_ns = 1000ll * atoll(c);  // test hook: forces the give-up path
        else D.wait_timeout_ns = ms * 1000000ll;
    }
    D.wind_rec = nullptr;
    {
        bool any_dyn = false;
        for (size_t i = 0; i < n_models; ++i) any_dyn |= res[i].dev.wind_dynamic != 0;
        if (any_dyn) {  // 16 B per row: which update() the wind getters of a row reflect (read-out kernels)
            FB_TRY(dev_alloc(e, &D.wind_rec, n_models * size_t(e->n_rows)));
            FB_TRYC(cudaMemset(D.wind_rec, 0, n_models * size_t(e->n_rows) * sizeof(WindRecord)));
        }
    }
    if (getenv("FB200_DEBUG"))
        fprintf(stderr, "[fb200] smem/CTA %zu B, CTAs/SM designed %d, resident %d, SMs %d\n", evolve_smem_bytes(), evolve_ctas_per_sm(),
                evolve_resident_ctas_per_sm(), e->sm_count);
    FB_TRY(dev_alloc(e, &D.cta_scratch, size_t(D.max_ctas) * (big ? fb200_big_evolve_scratch_doubles_per_cta() : evolve_scratch_doubles_per_cta())));
    if (D.star_ntri_max > 0) FB_TRY(dev_alloc(e, &D.star_T, size_t(D.max_ctas) * D.star_ntri_max));
    FB_TRY(dev_alloc(e, &D.prof, 16));
    FB_TRYC(cudaMemset(D.prof, 0, 16 * sizeof(long long)));
    dbg_lap("device allocations");
    {
        std::vector<fb200_model_dev_t> hm(n_models);
        std::vector<ModelState> hs(n_models);
        // 16 KB per model: packed by the host threads into buffers that are not value-initialised first (a single thread zeroing and
        // filling 200 MB for 12,500 models took 104 ms of a 170 ms create)
        std::unique_ptr<double[]> hh(new double[NN]), hF(new double[NN]);
        for_each_model([&](size_t i) {
            hm[i] = res[i].dev;
            ModelState& s = hs[i];
            std::memset(&s, 0, sizeof(s));
            s.t = res[i].dev.t0;
            s.F_in = res[i].dev.F_in0;
            s.Mdot_in_prev = -INFINITY;  // cpp/include/freddi_state.hpp:245
            s.first = res[i].dev.first0;
            s.last = Nx - 1;
            std::memcpy(&hh[i * Nx], res[i].h.data(), sizeof(double) * Nx);
            std::memcpy(&hF[i * Nx], res[i].F.data(), sizeof(double) * Nx);
        });
        dbg_lap("pack h, F, state");
        FB_TRYC(cudaMemcpy(d_models, hm.data(), sizeof(fb200_model_dev_t) * n_models, cudaMemcpyHostToDevice));
        FB_TRYC(cudaMemcpy(d_state, hs.data(), sizeof(ModelState) * n_models, cudaMemcpyHostToDevice));
        FB_TRYC(cudaMemcpy(d_h, hh.get(), sizeof(double) * NN, cudaMemcpyHostToDevice));
        FB_TRYC(cudaMemcpy(d_F, hF.get(), sizeof(double) * NN, cudaMemcpyHostToDevice));
        e->host_F0.swap(hF);
        e->host_state0.swap(hs);
        dbg_lap("H2D models, state, h, F");
    }
    FB_TRYC(cudaMemset(d_rows, 0xFF, rows_len * sizeof(double)));  // all-ones = NaN: rows never written stay NaN
    dbg_lap("memset rows");
    D.models = d_models; D.state = d_state; D.h = d_h; D.F = d_F; D.rows = d_rows;

    auto upload_opt = [&](bool any, const double** dst, auto getter) -> int {
        *dst = nullptr;
        if (!any) return FB200_OK;
        std::vector<double> host(NN, 0.);
        for (size_t i = 0; i < n_models; ++i) {
            const std::vector<double>& v = getter(res[i]);
            if (!v.empty()) std::memcpy(&host[i * Nx], v.data(), sizeof(double) * Nx);
        }
        double* p = nullptr;
        int rc = dev_alloc(e, &p, NN);
        if (rc != FB200_OK) return rc;
        cudaError_t ce = cudaMemcpy(p, host.data(), sizeof(double) * NN, cudaMemcpyHostToDevice);
        if (ce != cudaSuccess) return cuda_fail(ce, "cudaMemcpy");
        *dst = p;
        return FB200_OK;
    };
    FB_TRY(upload_opt(anyA, &D.wA, [](const Resolved& r) -> const std::vector<double>& { return r.wA; }));
    FB_TRY(upload_opt(anyB, &D.wB, [](const Resolved& r) -> const std::vector<double>& { return r.wB; }));
    if (anyC) {  // static wind C + d2Fmagn_dh2 (FreddiNeutronStarEvolution::windC, cpp/src/ns/ns_evolution.cpp:487-493)
        for (size_t i = 0; i < n_models; ++i) {
            Resolved& r = res[i];
            if (r.wC.empty()) r.wC.assign(Nx, 0.);
            if (!r.d2Fmagn_dh2.empty())
                for (int k = 0; k < Nx; ++k) r.wC[k] += r.d2Fmagn_dh2[k];
        }
    }
    FB_TRY(upload_opt(anyC, &D.wCs, [](const Resolved& r) -> const std::vector<double>& { return r.wC; }));
    FB_TRY(upload_opt(is_ns, &D.Fmagn, [](const Resolved& r) -> const std::vector<double>& { return r.Fmagn; }));
    FB_TRY(upload_opt(is_ns, &D.dFmagn_dh, [](const Resolved& r) -> const std::vector<double>& { return r.dFmagn_dh; }));
    FB_TRY(upload_opt(is_ns, &D.d2Fmagn_dh2, [](const Resolved& r) -> const std::vector<double>& { return r.d2Fmagn_dh2; }));
    D.Fsnap = nullptr;
    D.bounds = nullptr;
    if (cfg.record_F) {
        FB_TRY(dev_alloc(e, &D.Fsnap, n_models * size_t(e->n_rows) * Nx));
        FB_TRY(dev_alloc(e, &D.bounds, n_models * size_t(e->n_rows) * 2));
        FB_TRYC(cudaMemset(D.Fsnap, 0xFF, n_models * size_t(e->n_rows) * Nx * sizeof(double)));
        FB_TRYC(cudaMemset(D.bounds, 0xFF, n_models * size_t(e->n_rows) * 2 * sizeof(int)));
    }
    {
        double* p = nullptr;
        FB_TRY(readout_scratch_init(e, NN > n_models * 64 ? NN : n_models * 64, &p));
    }
#undef FB_TRY
#undef FB_TRYC
    dbg_lap("optional arrays, scratch");
    *out = e;
    return FB200_OK;
}

void fb200_ensemble_destroy(fb200_ensemble_t* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    for (void* p : e->allocs) cudaFree(p);
    if (e->scratch) cudaFree(e->scratch);
    if (e->pinned_F0) cudaFreeHost(e->pinned_F0);
    if (e->pinned_state0) cudaFreeHost(e->pinned_state0);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

size_t fb200_ensemble_n_models(const fb200_ensemble_t* e) { return e ? e->n : 0; }
size_t fb200_ensemble_n_cols(const fb200_ensemble_t* e) { return e ? size_t(e->n_cols) : 0; }
size_t fb200_ensemble_n_rows(const fb200_ensemble_t* e) { return e ? size_t(e->n_rows) : 0; }
size_t fb200_ensemble_nx(const fb200_ensemble_t* e) { return e ? size_t(e->Nx) : 0; }

int fb200_ensemble_info(const fb200_ensemble_t* e, fb200_model_info_t* out) {
    if (!e || !out) return set_err(FB200_ERR_BAD_HANDLE, "null handle");
    std::memcpy(out, e->info.data(), sizeof(fb200_model_info_t) * e->n);
    return FB200_OK;
}

int fb200_ensemble_reset(fb200_ensemble_t* e) {
    if (!e) return set_err(FB200_ERR_BAD_HANDLE, "null handle");
    int rc = select_device(e);
    if (rc != FB200_OK) return rc;
    const size_t NN = e->n * size_t(e->Nx);
    if (!e->pinned_F0) {
        FB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&e->pinned_F0), NN * sizeof(double)));
        FB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&e->pinned_state0), e->n * sizeof(ModelState)));
        std::memcpy(e->pinned_F0, e->host_F0.get(), NN * sizeof(double));
        std::memcpy(e->pinned_state0, e->host_state0.data(), e->n * sizeof(ModelState));
    }
    FB_CUDA(cudaMemcpyAsync(e->dev.F, e->pinned_F0, NN * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    FB_CUDA(cudaMemcpyAsync(e->dev.state, e->pinned_state0, e->n * sizeof(ModelState), cudaMemcpyHostToDevice, e->stream));
    FB_CUDA(cudaMemsetAsync(e->dev.rows, 0xFF, e->n * size_t(e->n_rows) * e->n_cols * sizeof(double), e->stream));
    FB_CUDA(cudaMemsetAsync(e->dev.failed, 0, e->n * sizeof(int), e->stream));
    if (e->dev.Fsnap) {
        FB_CUDA(cudaMemsetAsync(e->dev.Fsnap, 0xFF, e->n * size_t(e->n_rows) * e->Nx * sizeof(double), e->stream));
        FB_CUDA(cudaMemsetAsync(e->dev.bounds, 0xFF, e->n * size_t(e->n_rows) * 2 * sizeof(int), e->stream));
    }
    return FB200_OK;
}

int fb200_ensemble_evolve_async(fb200_ensemble_t* e, int64_t n_steps) {
    if (!e) return set_err(FB200_ERR_BAD_HANDLE, "null handle");
    int rc = select_device(e);
    if (rc != FB200_OK) return rc;
    FB_CUDA(cudaMemsetAsync(e->dev.queue, 0, sizeof(int), e->stream));
    FB_CUDA(cudaMemsetAsync(e->dev.progress, 0, sizeof(int) * e->n, e->stream));
    e->dev.n_steps = (n_steps < 0 || n_steps > 0x7fffffff) ? -1 : int(n_steps);
    {
        const long long most = (e->dev.n_steps < 0 || e->dev.n_steps > e->max_Nt) ? e->max_Nt : e->dev.n_steps;
        e->dev.chunk_steps = e->chunk_steps;
        e->dev.n_rounds = int(std::max<long long>(1, (most + e->chunk_steps - 1) / e->chunk_steps));
    }
    FB_CUDA(cudaEventRecord(e->ev0, e->stream));
    const int n_ctas = int(std::min<size_t>(e->n, size_t(e->dev.max_ctas)));
    cudaError_t ce = e->big ? fb200_big_launch_evolve(&e->dev, n_ctas, e->stream) : launch_evolve(e->dev, n_ctas, e->stream);
    if (ce != cudaSuccess) return cuda_fail(ce, "fb200_evolve_kernel launch");
    FB_CUDA(cudaEventRecord(e->ev1, e->stream));
    e->last_launches = 1;
    e->timed = true;
    return FB200_OK;
}

int fb200_ensemble_sync(fb200_ensemble_t* e) {
    if (!e) return set_err(FB200_ERR_BAD_HANDLE, "null handle");
    int rc = select_device(e);
    if (rc != FB200_OK) return rc;
    FB_CUDA(cudaStreamSynchronize(e->stream));
    return FB200_OK;
}

int fb200_ensemble_evolve(fb200_ensemble_t* e, int64_t n_steps) {
    int rc = fb200_ensemble_evolve_async(e, n_steps);
    if (rc != FB200_OK) return rc;
    return fb200_ensemble_sync(e);
}

int fb200_ensemble_last_evolve_ms(fb200_ensemble_t* e, float* ms) {
    if (!e || !ms) return set_err(FB200_ERR_BAD_HANDLE, "null handle");
    if (!e->timed) return set_err(FB200_ERR_LOGIC, "no evolve call has been made");
    int rc = select_device(e);
    if (rc != FB200_OK) return rc;
    FB_CUDA(cudaEventSynchronize(e->ev1));
    FB_CUDA(cudaEventElapsedTime(ms, e->ev0, e->ev1));
    return FB200_OK;
}

int fb200_ensemble_last_launches(const fb200_ensemble_t* e) { return e ? e->last_launches : 0; }

// Diagnostic (not part of include/freddi_b200.h): per-phase cycle sums of FB_PROFILE builds, zeros otherwise.
int fb200_debug_phase_cycles(fb200_ensemble_t* e, long long* out8) {
    if (!e || !out8) return FB200_ERR_BAD_HANDLE;
    if (select_device(e) != FB200_OK) return FB200_ERR_CUDA;
    cudaStreamSynchronize(e->stream);
    return cudaMemcpy(out8, e->dev.prof, 16 * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? FB200_OK : FB200_ERR_CUDA;
}

int fb200_ensemble_rows(fb200_ensemble_t* e, double* out, size_t out_len) {
    if (!e || !out) return set_err(FB200_ERR_BAD_HANDLE, "null handle");
    const size_t need = e->n * size_t(e->n_rows) * e->n_cols;
    if (out_len < need) return set_err(FB200_ERR_INVALID_ARGUMENT, "output buffer too small");
    int rc = select_device(e);
    if (rc != FB200_OK) return rc;
    FB_CUDA(cudaMemcpyAsync(out, e->dev.rows, need * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    FB_CUDA(cudaStreamSynchronize(e->stream));
    return FB200_OK;
}

int fb200_ensemble_rows_device(fb200_ensemble_t* e, const double** dptr) {
    if (!e || !dptr) return set_err(FB200_ERR_BAD_HANDLE, "null handle");
    *dptr = e->dev.rows;
    return FB200_OK;
}

static int fetch_state(fb200_ensemble_t* e, std::vector<ModelState>& hs) {
    int rc = select_device(e);
    if (rc != FB200_OK) return rc;
    hs.resize(e->n);
    std::vector<int> failed(e->n);
    FB_CUDA(cudaMemcpyAsync(hs.data(), e->dev.state, sizeof(ModelState) * e->n, cudaMemcpyDeviceToHost, e->stream));
    FB_CUDA(cudaMemcpyAsync(failed.data(), e->dev.failed, sizeof(int) * e->n, cudaMemcpyDeviceToHost, e->stream));
    FB_CUDA(cudaStreamSynchronize(e->stream));
    // a model whose work item gave up waiting (evolve.cu) keeps the rows its producers finished and is reported FAILED
    for (size_t i = 0; i < e->n; ++i)
        if (failed[i] && hs[i].status == FB200_MODEL_OK) hs[i].status = FB200_MODEL_FAILED;
    return FB200_OK;
}

int fb200_ensemble_status(fb200_ensemble_t* e, int32_t* status, int32_t* rows_valid, int64_t* picard_iters) {
    if (!e) return set_err(FB200_ERR_BAD_HANDLE, "null handle");
    std::vector<ModelState> hs;
    int rc = fetch_state(e, hs);
    if (rc != FB200_OK) return rc;
    for (size_t i = 0; i < e->n; ++i) {
        if (status) status[i] = hs[i].status;
        if (rows_valid) rows_valid[i] = hs[i].rows_valid;
        if (picard_iters) picard_iters[i] = hs[i].picard_iters;
    }
    return FB200_OK;
}

int fb200_ensemble_counters(fb200_ensemble_t* e, int64_t* picard_iters, int64_t* lx_trial_steps) {
    if (!e) return set_err(FB200_ERR_BAD_HANDLE, "null handle");
    std::vector<ModelState> hs;
    int rc = fetch_state(e, hs);
    if (rc != FB200_OK) return rc;
    for (size_t i = 0; i < e->n; ++i) {
        if (picard_iters) picard_iters[i] = hs[i].picard_iters;
        if (lx_trial_steps) lx_trial_steps[i] = hs[i].lx_trials;
    }
    return FB200_OK;
}

int fb200_ensemble_work_counters(fb200_ensemble_t* e, int64_t* out, size_t out_len) {
    if (!e || !out) return set_err(FB200_ERR_BAD_HANDLE, "null handle");
    if (out_len < e->n * size_t(FB200_N_WORK_COUNTERS)) return set_err(FB200_ERR_INVALID_ARGUMENT, "output buffer too small");
    std::vector<ModelState> hs;
    int rc = fetch_state(e, hs);
    if (rc != FB200_OK) return rc;
    for (size_t i = 0; i < e->n; ++i) {
        int64_t* o = out + i * FB200_N_WORK_COUNTERS;
        o[FB200_WORK_PICARD_ITERS] = hs[i].picard_iters;
        o[FB200_WORK_LX_TRIALS] = hs[i].lx_trials;
        o[FB200_WORK_RING_STEPS] = hs[i].ring_steps;
        o[FB200_WORK_RING_ITERS] = hs[i].ring_iters;
        o[FB200_WORK_POW_EVALS] = hs[i].pow_evals;
        o[FB200_WORK_ROW_RINGS] = hs[i].row_rings;
    }
    return FB200_OK;
}

int fb200_ensemble_state(fb200_ensemble_t* e, double* t, int64_t* i_t, int64_t* first, int64_t* last) {
    if (!e) return set_err(FB200_ERR_BAD_HANDLE, "null handle");
    std::vector<ModelState> hs;
    int rc = fetch_state(e, hs);
    if (rc != FB200_OK) return rc;
    for (size_t i = 0; i < e->n; ++i) {
        if (t) t[i] = hs[i].t;
        if (i_t) i_t[i] = hs[i].i_t;
        if (first) first[i] = hs[i].first;
        if (last) last[i] = hs[i].last;
    }
    return FB200_OK;
}

int fb200_ensemble_row_bounds(fb200_ensemble_t* e, int64_t row, int64_t* first, int64_t* last) {
    if (!e) return set_err(FB200_ERR_BAD_HANDLE, "null handle");
    if (!e->dev.bounds) return set_err(FB200_ERR_LOGIC, "row bounds need an ensemble created with record_F");
    if (row < 0 || row >= e->n_rows) return set_err(FB200_ERR_INVALID_ARGUMENT, "row out of range");
    int rc = select_device(e);
    if (rc != FB200_OK) return rc;
    std::vector<int> hb(e->n * size_t(e->n_rows) * 2);
    FB_CUDA(cudaMemcpyAsync(hb.data(), e->dev.bounds, hb.size() * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    FB_CUDA(cudaStreamSynchronize(e->stream));
    for (size_t i = 0; i < e->n; ++i) {
        const int* b = &hb[(i * e->n_rows + size_t(row)) * 2];
        if (first) first[i] = b[0];
        if (last) last[i] = b[1];
    }
    return FB200_OK;
}

// The read-out staging buffer of the ensemble, at least `need` doubles: allocated once, grown when a larger request
// comes, freed by fb200_ensemble_destroy (no cudaMalloc / cudaFree per call).
static int readout_scratch(fb200_ensemble_t* e, size_t need, double** p) {
    if (need > e->scratch_len) {
        FB_CUDA(cudaStreamSynchronize(e->stream));
        if (e->scratch) FB_CUDA(cudaFree(e->scratch));
        e->scratch = nullptr;
        e->scratch_len = 0;
        const size_t len = need + need / 4;
        FB_CUDA(cudaMalloc(reinterpret_cast<void**>(&e->scratch), len * sizeof(double)));
        e->scratch_len = len;
    }
    *p = e->scratch;
    return FB200_OK;
}

static int check_row(fb200_ensemble_t* e, int64_t row) {
    if (row >= 0) {
        if (!e->dev.Fsnap) return set_err(FB200_ERR_LOGIC, "reading a recorded row needs an ensemble created with record_F");
        if (row >= e->n_rows) return set_err(FB200_ERR_INVALID_ARGUMENT, "row out of range");
    }
    return FB200_OK;
}

// field of rows row0 .. row0 + n_sel - 1 of every model -> host out[model][n_sel][Nx]
static int field_rows(fb200_ensemble_t* e, int field, int64_t row0, int n_sel, int pad_nan, double* out, size_t out_len) {
    if (!e || !out) return set_err(FB200_ERR_BAD_HANDLE, "null handle");
    if (field < 0 || field >= FB200_N_FIELDS) return set_err(FB200_ERR_INVALID_ARGUMENT, "unknown field");
    if (n_sel < 1) return set_err(FB200_ERR_INVALID_ARGUMENT, "n_rows must be positive");
    const size_t need = e->n * size_t(n_sel) * size_t(e->Nx);
    if (out_len < need) return set_err(FB200_ERR_INVALID_ARGUMENT, "output buffer too small");
    int rc = check_row(e, row0);
    if (rc != FB200_OK) return rc;
    if (row0 >= 0 && row0 + n_sel > e->n_rows) return set_err(FB200_ERR_INVALID_ARGUMENT, "row range out of bounds");
    if (row0 < 0 && n_sel != 1) return set_err(FB200_ERR_INVALID_ARGUMENT, "the current state is a single row");
    rc = select_device(e);
    if (rc != FB200_OK) return rc;
    double* d_o = nullptr;
    rc = readout_scratch(e, need, &d_o);
    if (rc != FB200_OK) return rc;
    cudaError_t ce = e->big ? fb200_big_launch_field(&e->dev, field, row0, n_sel, pad_nan, d_o, e->stream)
                            : launch_field(e->dev, field, row0, n_sel, pad_nan, d_o, e->stream);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(out, d_o, need * sizeof(double), cudaMemcpyDeviceToHost, e->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
    if (ce != cudaSuccess) return cuda_fail(ce, "fb200_field_kernel");
    return FB200_OK;
}

int fb200_ensemble_field(fb200_ensemble_t* e, int field, int64_t row, double* out, size_t out_len) {
    return field_rows(e, field, row, 1, 0, out, out_len);
}

int fb200_ensemble_field_rows(fb200_ensemble_t* e, int field, int64_t row_begin, int64_t n_rows, int pad_nan, double* out, size_t out_len) {
    if (row_begin < 0) return set_err(FB200_ERR_INVALID_ARGUMENT, "row_begin must be a recorded row");
    if (n_rows < 1 || n_rows > 0x7fffffff) return set_err(FB200_ERR_INVALID_ARGUMENT, "n_rows out of range");
    return field_rows(e, field, row_begin, int(n_rows), pad_nan, out, out_len);
}

// region 0 hot, 1 cold, 2 companion star (needs phases[n_sel] and an ensemble created with star_flux)
static int flux_rows(fb200_ensemble_t* e, int region, int64_t row0, int n_sel, const double* lambdas, size_t n_lambdas,
                     const double* phases, double* out, size_t out_len) {
    if (!e || !out || !lambdas) return set_err(FB200_ERR_BAD_HANDLE, "null handle");
    if (region < 0 || region > 2) return set_err(FB200_ERR_INVALID_ARGUMENT, "region must be 0 (hot), 1 (cold) or 2 (star)");
    if (region == 2 && (!e->dev.cfg.star_data || !phases))
        return set_err(FB200_ERR_LOGIC, "the star flux needs an ensemble created with star_flux and a phase per row");
    if (n_lambdas == 0) return FB200_OK;
    if (n_sel < 1) return set_err(FB200_ERR_INVALID_ARGUMENT, "n_rows must be positive");
    const size_t need = e->n * size_t(n_sel) * n_lambdas;
    if (out_len < need) return set_err(FB200_ERR_INVALID_ARGUMENT, "output buffer too small");
    int rc = check_row(e, row0);
    if (rc != FB200_OK) return rc;
    if (row0 >= 0 && row0 + n_sel > e->n_rows) return set_err(FB200_ERR_INVALID_ARGUMENT, "row range out of bounds");
    if (row0 < 0 && n_sel != 1) return set_err(FB200_ERR_INVALID_ARGUMENT, "the current state is a single row");
    rc = select_device(e);
    if (rc != FB200_OK) return rc;
    // one persistent scratch buffer (grown on demand, freed with the ensemble): wavelengths, phases, then the result
    double *d_l = nullptr, *d_o = nullptr, *d_p = nullptr;
    rc = readout_scratch(e, n_lambdas + size_t(n_sel) + need, &d_l);
    if (rc != FB200_OK) return rc;
    d_p = d_l + n_lambdas;
    d_o = d_p + n_sel;
    cudaError_t ce = cudaMemcpyAsync(d_l, lambdas, n_lambdas * sizeof(double), cudaMemcpyHostToDevice, e->stream);
    if (ce == cudaSuccess && phases) ce = cudaMemcpyAsync(d_p, phases, size_t(n_sel) * sizeof(double), cudaMemcpyHostToDevice, e->stream);
    if (ce == cudaSuccess)
        ce = e->big ? fb200_big_launch_flux(&e->dev, region, row0, n_sel, d_l, int(n_lambdas), d_p, d_o, e->stream)
                    : launch_flux(e->dev, region, row0, n_sel, d_l, int(n_lambdas), d_p, d_o, e->stream);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(out, d_o, need * sizeof(double), cudaMemcpyDeviceToHost, e->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
    if (ce != cudaSuccess) return cuda_fail(ce, "fb200_flux_kernel");
    return FB200_OK;
}

int fb200_ensemble_flux(fb200_ensemble_t* e, int region, int64_t row, const double* lambdas, size_t n_lambdas, double* out,
                        size_t out_len) {
    if (region == 2) return set_err(FB200_ERR_INVALID_ARGUMENT, "use fb200_ensemble_flux_rows for the star (it needs a phase)");
    return flux_rows(e, region, row, 1, lambdas, n_lambdas, nullptr, out, out_len);
}

int fb200_ensemble_flux_rows(fb200_ensemble_t* e, int region, int64_t row_begin, int64_t n_rows, const double* lambdas, size_t n_lambdas,
                             const double* phases, double* out, size_t out_len) {
    if (n_rows < 1 || n_rows > 0x7fffffff) return set_err(FB200_ERR_INVALID_ARGUMENT, "n_rows out of range");
    if (row_begin < 0 && n_rows != 1) return set_err(FB200_ERR_INVALID_ARGUMENT, "the current state is a single row");
    return flux_rows(e, region, row_begin, int(n_rows), lambdas, n_lambdas, phases, out, out_len);
}

int fb200_ensemble_mdot_wind(fb200_ensemble_t* e, int64_t row, double* out, size_t out_len) {
    if (!e || !out) return set_err(FB200_ERR_BAD_HANDLE, "null handle");
    if (out_len < e->n) return set_err(FB200_ERR_INVALID_ARGUMENT, "output buffer too small");
    int rc = check_row(e, row);
    if (rc != FB200_OK) return rc;
    rc = select_device(e);
    if (rc != FB200_OK) return rc;
    cudaError_t ce = e->big ? fb200_big_launch_mdot_wind(&e->dev, row, e->scratch, e->stream)
                            : launch_mdot_wind(e->dev, row, e->scratch, e->stream);
    if (ce != cudaSuccess) return cuda_fail(ce, "fb200_mdot_wind_kernel launch");
    FB_CUDA(cudaMemcpyAsync(out, e->scratch, e->n * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    FB_CUDA(cudaStreamSynchronize(e->stream));
    return FB200_OK;
}

}  // extern "C"
