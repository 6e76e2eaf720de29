// extern "C" layer of include/freddi_b200.h: owns device memory, resolves arguments on the host, launches
// the kernels.  No CPU implementation of the path lives here: without a CUDA device every compute call fails.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "ensemble_impl.hpp"

using namespace fb200;

namespace {
int set_err(int code, const std::string& msg) { return fb_set_err(code, msg); }
int cuda_fail(cudaError_t e, const char* what) { return fb_cuda_fail(e, what); }

int select_device(const fb200_ensemble* e) {
    FB_CUDA(cudaSetDevice(e->device));
    return FB200_OK;
}
}  // namespace

extern "C" {

int fb200_abi_version(void) { return FB200_ABI_VERSION; }
const char* fb200_last_error(void) { return fb_err_msg().c_str(); }

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
    fb200_config_t cfg;
    if (cfg_in) cfg = *cfg_in; else fb200_config_default(&cfg);
    const int device = cfg.device;
    return build_ensembles(args, n_models, &cfg, &device, 1, /*keep_initial=*/true, nullptr, out, bad_model);
}

void fb200_ensemble_destroy(fb200_ensemble_t* e) { ensemble_free(e); }

void fb200_trim(void) { caches_trim(); }

int fb200_host_alloc(size_t bytes, void** p) {
    if (!p) return set_err(FB200_ERR_INVALID_ARGUMENT, "null argument");
    std::unique_lock<std::shared_mutex> no_open_window(g_upload_window);
    FB_CUDA(cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocPortable));
    return FB200_OK;
}

void fb200_host_free(void* p) {
    if (!p) return;
    std::unique_lock<std::shared_mutex> no_open_window(g_upload_window);
    cudaFreeHost(p);
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
    if (!e->F0_dev) return set_err(FB200_ERR_LOGIC, "this ensemble keeps no copy of its initial state");
    FB_CUDA(cudaMemcpyAsync(e->dev.F, e->F0_dev, NN * sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
    FB_CUDA(cudaMemcpyAsync(e->dev.state, e->state0_dev, e->n * sizeof(ModelState), cudaMemcpyDeviceToDevice, e->stream));
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
    return ensemble_launch(e, n_steps);
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

int fb200_ensemble_device_ptr(fb200_ensemble_t* e, int which, const void** dptr, int32_t* device) {
    if (!e || !dptr) return set_err(FB200_ERR_BAD_HANDLE, "null handle");
    const void* p = nullptr;
    switch (which) {
        case FB200_DEVPTR_ROWS: p = e->dev.rows; break;
        case FB200_DEVPTR_F: p = e->dev.F; break;
        case FB200_DEVPTR_H: p = e->dev.h; break;
        case FB200_DEVPTR_FSNAP: p = e->dev.Fsnap; break;
        case FB200_DEVPTR_BOUNDS: p = e->dev.bounds; break;
        default: return set_err(FB200_ERR_INVALID_ARGUMENT, "unknown device array");
    }
    if (!p) return set_err(FB200_ERR_LOGIC, "this array needs an ensemble created with record_F");
    *dptr = p;
    if (device) *device = e->device;
    return FB200_OK;
}

// ---- checkpoint / resume: the copyable state of the reference (FreddiState's copy constructor, cpp/src/freddi_state.cpp:84-91) ----
namespace {
struct CheckpointHeader {
    unsigned long long magic;
    unsigned long long n, Nx, n_rows, n_cols, with_rows, state_bytes;
};
constexpr unsigned long long kCheckpointMagic = 0x4642323030434b31ull;  // "FB200CK1"
}  // namespace

size_t fb200_ensemble_checkpoint_bytes(const fb200_ensemble_t* e, int with_rows) {
    if (!e) return 0;
    const size_t NN = e->n * size_t(e->Nx);
    size_t b = sizeof(CheckpointHeader) + e->n * sizeof(ModelState) + e->n * sizeof(int) + 2 * NN * sizeof(double);
    if (with_rows) b += e->n * size_t(e->n_rows) * e->n_cols * sizeof(double);
    if (with_rows && e->dev.wind_rec) b += e->n * size_t(e->n_rows) * sizeof(WindRecord);  // what the wind getters of every row reflect
    return b;
}

int fb200_ensemble_checkpoint_save(fb200_ensemble_t* e, void* buf, size_t len, int with_rows) {
    if (!e || !buf) return set_err(FB200_ERR_BAD_HANDLE, "null handle");
    if (len < fb200_ensemble_checkpoint_bytes(e, with_rows)) return set_err(FB200_ERR_INVALID_ARGUMENT, "checkpoint buffer too small");
    int rc = select_device(e);
    if (rc != FB200_OK) return rc;
    const size_t NN = e->n * size_t(e->Nx);
    char* p = static_cast<char*>(buf);
    CheckpointHeader h{kCheckpointMagic, e->n, (unsigned long long)e->Nx, (unsigned long long)e->n_rows, (unsigned long long)e->n_cols,
                       with_rows ? 1ull : 0ull, sizeof(ModelState)};
    std::memcpy(p, &h, sizeof(h));
    p += sizeof(h);
    FB_CUDA(cudaMemcpyAsync(p, e->dev.state, e->n * sizeof(ModelState), cudaMemcpyDeviceToHost, e->stream));
    p += e->n * sizeof(ModelState);
    FB_CUDA(cudaMemcpyAsync(p, e->dev.failed, e->n * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    p += e->n * sizeof(int);
    FB_CUDA(cudaMemcpyAsync(p, e->dev.F, NN * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    p += NN * sizeof(double);
    FB_CUDA(cudaMemcpyAsync(p, e->dev.K, NN * sizeof(double), cudaMemcpyDeviceToHost, e->stream));  // the closure carried with F (k_valid)
    p += NN * sizeof(double);
    if (with_rows) {
        const size_t rb = e->n * size_t(e->n_rows) * e->n_cols * sizeof(double);
        FB_CUDA(cudaMemcpyAsync(p, e->dev.rows, rb, cudaMemcpyDeviceToHost, e->stream));
        p += rb;
        if (e->dev.wind_rec) FB_CUDA(cudaMemcpyAsync(p, e->dev.wind_rec, e->n * size_t(e->n_rows) * sizeof(WindRecord), cudaMemcpyDeviceToHost, e->stream));
    }
    FB_CUDA(cudaStreamSynchronize(e->stream));
    return FB200_OK;
}

int fb200_ensemble_checkpoint_load(fb200_ensemble_t* e, const void* buf, size_t len) {
    if (!e || !buf) return set_err(FB200_ERR_BAD_HANDLE, "null handle");
    if (len < sizeof(CheckpointHeader)) return set_err(FB200_ERR_INVALID_ARGUMENT, "not a checkpoint");
    CheckpointHeader h;
    std::memcpy(&h, buf, sizeof(h));
    if (h.magic != kCheckpointMagic || h.state_bytes != sizeof(ModelState)) return set_err(FB200_ERR_INVALID_ARGUMENT, "not a checkpoint of this library version");
    if (h.n != e->n || h.Nx != (unsigned long long)e->Nx || h.n_rows != (unsigned long long)e->n_rows || h.n_cols != (unsigned long long)e->n_cols)
        return set_err(FB200_ERR_INVALID_ARGUMENT, "the checkpoint belongs to an ensemble of another shape");
    if (len < fb200_ensemble_checkpoint_bytes(e, int(h.with_rows))) return set_err(FB200_ERR_INVALID_ARGUMENT, "truncated checkpoint");
    int rc = select_device(e);
    if (rc != FB200_OK) return rc;
    const size_t NN = e->n * size_t(e->Nx);
    const char* p = static_cast<const char*>(buf) + sizeof(h);
    FB_CUDA(cudaMemcpyAsync(e->dev.state, p, e->n * sizeof(ModelState), cudaMemcpyHostToDevice, e->stream));
    p += e->n * sizeof(ModelState);
    FB_CUDA(cudaMemcpyAsync(e->dev.failed, p, e->n * sizeof(int), cudaMemcpyHostToDevice, e->stream));
    p += e->n * sizeof(int);
    FB_CUDA(cudaMemcpyAsync(e->dev.F, p, NN * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    p += NN * sizeof(double);
    FB_CUDA(cudaMemcpyAsync(e->dev.K, p, NN * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    p += NN * sizeof(double);
    if (h.with_rows) {
        const size_t rb = e->n * size_t(e->n_rows) * e->n_cols * sizeof(double);
        FB_CUDA(cudaMemcpyAsync(e->dev.rows, p, rb, cudaMemcpyHostToDevice, e->stream));
        p += rb;
        if (e->dev.wind_rec) FB_CUDA(cudaMemcpyAsync(e->dev.wind_rec, p, e->n * size_t(e->n_rows) * sizeof(WindRecord), cudaMemcpyHostToDevice, e->stream));
    }
    FB_CUDA(cudaStreamSynchronize(e->stream));  // the caller's buffer may go away
    return FB200_OK;
}

int fb200_ensemble_get_state(fb200_ensemble_t* e, double* F, double* t, int64_t* i_t, int64_t* first, int64_t* last, double* F_in,
                             double* Mdot_in_prev) {
    if (!e) return set_err(FB200_ERR_BAD_HANDLE, "null handle");
    int rc = select_device(e);
    if (rc != FB200_OK) return rc;
    std::vector<ModelState> hs(e->n);
    FB_CUDA(cudaMemcpyAsync(hs.data(), e->dev.state, sizeof(ModelState) * e->n, cudaMemcpyDeviceToHost, e->stream));
    if (F) FB_CUDA(cudaMemcpyAsync(F, e->dev.F, e->n * size_t(e->Nx) * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    FB_CUDA(cudaStreamSynchronize(e->stream));
    for (size_t i = 0; i < e->n; ++i) {
        if (t) t[i] = hs[i].t;
        if (i_t) i_t[i] = hs[i].i_t;
        if (first) first[i] = hs[i].first;
        if (last) last[i] = hs[i].last;
        if (F_in) F_in[i] = hs[i].F_in;
        if (Mdot_in_prev) Mdot_in_prev[i] = hs[i].Mdot_in_prev;
    }
    return FB200_OK;
}

int fb200_ensemble_set_state(fb200_ensemble_t* e, const double* F, const double* t, const int64_t* i_t, const int64_t* first,
                             const int64_t* last, const double* F_in, const double* Mdot_in_prev) {
    if (!e) return set_err(FB200_ERR_BAD_HANDLE, "null handle");
    int rc = select_device(e);
    if (rc != FB200_OK) return rc;
    std::vector<ModelState> hs(e->n);
    FB_CUDA(cudaMemcpyAsync(hs.data(), e->dev.state, sizeof(ModelState) * e->n, cudaMemcpyDeviceToHost, e->stream));
    FB_CUDA(cudaStreamSynchronize(e->stream));
    for (size_t i = 0; i < e->n; ++i) {
        ModelState& s = hs[i];
        if (t) s.t = t[i];
        if (F_in) s.F_in = F_in[i];
        if (Mdot_in_prev) s.Mdot_in_prev = Mdot_in_prev[i];
        if (first) s.first = int(first[i]);
        if (last) s.last = int(last[i]);
        if (i_t) s.i_t = int(i_t[i]);
        if (s.first < 0 || s.last >= e->Nx || s.first + 2 > s.last || s.i_t < 0 || s.i_t >= e->n_rows)
            return set_err(FB200_ERR_INVALID_ARGUMENT, "state of model " + std::to_string(i) + ": need 0 <= first, first + 2 <= last < Nx, 0 <= i_t <= Nt");
        // the caller's state replaces the evolved one: the model is alive again, its next row is i_t + 1 (row 0 is written from
        // this state when i_t == 0), and the closure K of the old F is void
        s.status = FB200_MODEL_OK;
        s.rows_valid = s.i_t + (s.i_t > 0 ? 1 : 0);
        s.row0_done = s.i_t > 0 ? 1 : 0;
        s.k_valid = 0;
    }
    FB_CUDA(cudaMemcpyAsync(e->dev.state, hs.data(), sizeof(ModelState) * e->n, cudaMemcpyHostToDevice, e->stream));
    if (F) FB_CUDA(cudaMemcpyAsync(e->dev.F, F, e->n * size_t(e->Nx) * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    FB_CUDA(cudaMemsetAsync(e->dev.failed, 0, e->n * sizeof(int), e->stream));
    FB_CUDA(cudaStreamSynchronize(e->stream));
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
        arena_release(e->scratch, e->stream);
        e->scratch = nullptr;
        e->scratch_len = 0;
        const size_t len = need + need / 4;
        void* q = nullptr;
        const int rc = arena_acquire(e->device, len * sizeof(double), e->stream, &q);
        if (rc != FB200_OK) return rc;
        e->scratch = static_cast<double*>(q);
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
    cudaError_t ce = e->kb->launch_field(&e->dev, field, row0, n_sel, pad_nan, d_o, e->stream);
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
        ce = e->kb->launch_flux(&e->dev, region, row0, n_sel, d_l, int(n_lambdas), d_p, d_o, e->stream);
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
    double* d_o = nullptr;
    rc = readout_scratch(e, e->n, &d_o);
    if (rc != FB200_OK) return rc;
    cudaError_t ce = e->kb->launch_mdot_wind(&e->dev, row, d_o, e->stream);
    if (ce != cudaSuccess) return cuda_fail(ce, "fb200_mdot_wind_kernel launch");
    FB_CUDA(cudaMemcpyAsync(out, d_o, e->n * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    FB_CUDA(cudaStreamSynchronize(e->stream));
    return FB200_OK;
}

}  // extern "C"
