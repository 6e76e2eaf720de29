// In-process multi-GPU host (include/freddi_b200.h, "sharded" and "sweep" entry points): one ensemble per device behind one
// handle, models dealt out by interleaved index (model k -> shard k mod G, local index k / G) so that cost correlated with the
// sweep parameters is balanced; no collective on the data path — the shards never exchange anything.  Every device call is
// asynchronous, so one host thread drives all devices; results land in disjoint strided slices of ONE caller buffer.
// Stands in for running the reference's driver (cpp/main.cpp:7-12 -> cpp/include/application.hpp:17-43) once per model.
#include <algorithm>
#include <cstring>
#include <vector>

#include "ensemble_impl.hpp"

using namespace fb200;

struct fb200_sharded {
    std::vector<fb200_ensemble*> shard;
    size_t n = 0;
    int n_rows = 0, n_cols = 0, Nx = 0;
};

namespace {

int resolve_devices(const int32_t* devices, int32_t n_devices, std::vector<int>& out) {
    out.clear();
    if (devices && n_devices > 0) {
        out.assign(devices, devices + n_devices);
        return FB200_OK;
    }
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        const std::string why = (ce != cudaSuccess) ? std::string("cudaGetDeviceCount: ") + cudaGetErrorString(ce) : std::string("no CUDA device");
        cudaGetLastError();
        return fb_set_err(FB200_ERR_NO_DEVICE, "no CUDA device: " + why + " (there is no CPU fallback)");
    }
    for (int d = 0; d < ndev; ++d) out.push_back(d);
    return FB200_OK;
}

struct DeviceGuard {  // the caller's current device is left as it was
    int dev = -1;
    DeviceGuard() { if (cudaGetDevice(&dev) != cudaSuccess) { dev = -1; cudaGetLastError(); } }
    ~DeviceGuard() { if (dev >= 0) cudaSetDevice(dev); }
};

// per-shard call into a temporary, scattered to the caller's interleaved order
template <class T, class Call>
int gather(fb200_sharded* h, T* out, size_t per_model, Call&& call) {
    const size_t G = h->shard.size();
    std::vector<T> tmp;
    for (size_t d = 0; d < G; ++d) {
        fb200_ensemble* e = h->shard[d];
        tmp.resize(e->n * per_model);
        const int rc = call(e, tmp.data(), tmp.size());
        if (rc != FB200_OK) return rc;
        for (size_t j = 0; j < e->n; ++j) std::memcpy(out + (d + j * G) * per_model, tmp.data() + j * per_model, per_model * sizeof(T));
    }
    return FB200_OK;
}

}  // namespace

extern "C" {

int fb200_sharded_create(const fb200_args_t* args, size_t n_models, const fb200_config_t* cfg, const int32_t* devices, int32_t n_devices,
                         fb200_sharded_t** out, size_t* bad_model) {
    if (bad_model) *bad_model = 0;
    if (!args || !out || n_models == 0) return fb_set_err(FB200_ERR_INVALID_ARGUMENT, "null argument or empty ensemble");
    std::vector<int> devs;
    int rc = resolve_devices(devices, n_devices, devs);
    if (rc != FB200_OK) return rc;
    if (devs.size() > n_models) devs.resize(n_models);
    std::vector<fb200_ensemble*> shards(devs.size(), nullptr);
    rc = build_ensembles(args, n_models, cfg, devs.data(), int(devs.size()), /*keep_initial=*/true, nullptr, shards.data(), bad_model);
    if (rc != FB200_OK) return rc;
    auto* h = new fb200_sharded();
    h->shard = shards;
    h->n = n_models;
    h->n_rows = shards[0]->n_rows;
    h->n_cols = shards[0]->n_cols;
    h->Nx = shards[0]->Nx;
    *out = h;
    return FB200_OK;
}

void fb200_sharded_destroy(fb200_sharded_t* h) {
    if (!h) return;
    DeviceGuard guard;
    for (fb200_ensemble* e : h->shard) ensemble_free(e);
    delete h;
}

int32_t fb200_sharded_n_shards(const fb200_sharded_t* h) { return h ? int32_t(h->shard.size()) : 0; }
fb200_ensemble_t* fb200_sharded_shard(fb200_sharded_t* h, int32_t k) {
    return (h && k >= 0 && size_t(k) < h->shard.size()) ? h->shard[size_t(k)] : nullptr;
}
size_t fb200_sharded_n_models(const fb200_sharded_t* h) { return h ? h->n : 0; }
size_t fb200_sharded_n_rows(const fb200_sharded_t* h) { return h ? size_t(h->n_rows) : 0; }
size_t fb200_sharded_n_cols(const fb200_sharded_t* h) { return h ? size_t(h->n_cols) : 0; }
size_t fb200_sharded_nx(const fb200_sharded_t* h) { return h ? size_t(h->Nx) : 0; }

int fb200_sharded_reset(fb200_sharded_t* h) {
    if (!h) return fb_set_err(FB200_ERR_BAD_HANDLE, "null handle");
    DeviceGuard guard;
    for (fb200_ensemble* e : h->shard) {
        const int rc = fb200_ensemble_reset(e);
        if (rc != FB200_OK) return rc;
    }
    return FB200_OK;
}

int fb200_sharded_evolve_async(fb200_sharded_t* h, int64_t n_steps) {
    if (!h) return fb_set_err(FB200_ERR_BAD_HANDLE, "null handle");
    DeviceGuard guard;
    for (fb200_ensemble* e : h->shard) {  // every launch is asynchronous: all devices run at once
        const int rc = fb200_ensemble_evolve_async(e, n_steps);
        if (rc != FB200_OK) return rc;
    }
    return FB200_OK;
}

int fb200_sharded_sync(fb200_sharded_t* h) {
    if (!h) return fb_set_err(FB200_ERR_BAD_HANDLE, "null handle");
    DeviceGuard guard;
    for (fb200_ensemble* e : h->shard) {
        const int rc = fb200_ensemble_sync(e);
        if (rc != FB200_OK) return rc;
    }
    return FB200_OK;
}

int fb200_sharded_evolve(fb200_sharded_t* h, int64_t n_steps) {
    const int rc = fb200_sharded_evolve_async(h, n_steps);
    return rc != FB200_OK ? rc : fb200_sharded_sync(h);
}

int fb200_sharded_last_evolve_ms(fb200_sharded_t* h, float* ms) {
    if (!h || !ms) return fb_set_err(FB200_ERR_BAD_HANDLE, "null handle");
    DeviceGuard guard;
    float most = 0.f;
    for (fb200_ensemble* e : h->shard) {
        float t = 0.f;
        const int rc = fb200_ensemble_last_evolve_ms(e, &t);
        if (rc != FB200_OK) return rc;
        most = std::max(most, t);
    }
    *ms = most;
    return FB200_OK;
}

int fb200_sharded_rows(fb200_sharded_t* h, double* out, size_t out_len) {
    if (!h || !out) return fb_set_err(FB200_ERR_BAD_HANDLE, "null handle");
    const size_t row_len = size_t(h->n_rows) * h->n_cols, row_bytes = row_len * sizeof(double);
    if (out_len < h->n * row_len) return fb_set_err(FB200_ERR_INVALID_ARGUMENT, "output buffer too small");
    DeviceGuard guard;
    const size_t G = h->shard.size();
    for (size_t d = 0; d < G; ++d) {  // strided slices of the one caller buffer, all devices copying at once
        fb200_ensemble* e = h->shard[d];
        FB_CUDA(cudaSetDevice(e->device));
        if (G == 1) FB_CUDA(cudaMemcpyAsync(out, e->dev.rows, e->n * row_bytes, cudaMemcpyDeviceToHost, e->stream));
        else FB_CUDA(cudaMemcpy2DAsync(out + d * row_len, G * row_bytes, e->dev.rows, row_bytes, row_bytes, e->n, cudaMemcpyDeviceToHost, e->stream));
    }
    for (fb200_ensemble* e : h->shard) {
        FB_CUDA(cudaSetDevice(e->device));
        FB_CUDA(cudaStreamSynchronize(e->stream));
    }
    return FB200_OK;
}

int fb200_sharded_status(fb200_sharded_t* h, int32_t* status, int32_t* rows_valid, int64_t* picard_iters) {
    if (!h) return fb_set_err(FB200_ERR_BAD_HANDLE, "null handle");
    DeviceGuard guard;
    const size_t G = h->shard.size();
    std::vector<int32_t> st, rv;
    std::vector<int64_t> pi;
    for (size_t d = 0; d < G; ++d) {
        fb200_ensemble* e = h->shard[d];
        st.resize(e->n); rv.resize(e->n); pi.resize(e->n);
        const int rc = fb200_ensemble_status(e, st.data(), rv.data(), pi.data());
        if (rc != FB200_OK) return rc;
        for (size_t j = 0; j < e->n; ++j) {
            if (status) status[d + j * G] = st[j];
            if (rows_valid) rows_valid[d + j * G] = rv[j];
            if (picard_iters) picard_iters[d + j * G] = pi[j];
        }
    }
    return FB200_OK;
}

int fb200_sharded_state(fb200_sharded_t* h, double* t, int64_t* i_t, int64_t* first, int64_t* last) {
    if (!h) return fb_set_err(FB200_ERR_BAD_HANDLE, "null handle");
    DeviceGuard guard;
    const size_t G = h->shard.size();
    std::vector<double> tt;
    std::vector<int64_t> a, b, c;
    for (size_t d = 0; d < G; ++d) {
        fb200_ensemble* e = h->shard[d];
        tt.resize(e->n); a.resize(e->n); b.resize(e->n); c.resize(e->n);
        const int rc = fb200_ensemble_state(e, tt.data(), a.data(), b.data(), c.data());
        if (rc != FB200_OK) return rc;
        for (size_t j = 0; j < e->n; ++j) {
            if (t) t[d + j * G] = tt[j];
            if (i_t) i_t[d + j * G] = a[j];
            if (first) first[d + j * G] = b[j];
            if (last) last[d + j * G] = c[j];
        }
    }
    return FB200_OK;
}

int fb200_sharded_info(fb200_sharded_t* h, fb200_model_info_t* out) {
    if (!h || !out) return fb_set_err(FB200_ERR_BAD_HANDLE, "null handle");
    const size_t G = h->shard.size();
    for (size_t d = 0; d < G; ++d)
        for (size_t j = 0; j < h->shard[d]->n; ++j) out[d + j * G] = h->shard[d]->info[j];
    return FB200_OK;
}

int fb200_sharded_work_counters(fb200_sharded_t* h, int64_t* out, size_t out_len) {
    if (!h || !out) return fb_set_err(FB200_ERR_BAD_HANDLE, "null handle");
    if (out_len < h->n * size_t(FB200_N_WORK_COUNTERS)) return fb_set_err(FB200_ERR_INVALID_ARGUMENT, "output buffer too small");
    DeviceGuard guard;
    return gather<int64_t>(h, out, FB200_N_WORK_COUNTERS, [](fb200_ensemble* e, int64_t* p, size_t len) { return fb200_ensemble_work_counters(e, p, len); });
}

int fb200_sharded_field(fb200_sharded_t* h, int field, int64_t row, double* out, size_t out_len) {
    if (!h || !out) return fb_set_err(FB200_ERR_BAD_HANDLE, "null handle");
    if (out_len < h->n * size_t(h->Nx)) return fb_set_err(FB200_ERR_INVALID_ARGUMENT, "output buffer too small");
    DeviceGuard guard;
    return gather<double>(h, out, size_t(h->Nx), [&](fb200_ensemble* e, double* p, size_t len) { return fb200_ensemble_field(e, field, row, p, len); });
}

int fb200_sharded_flux(fb200_sharded_t* h, int region, int64_t row, const double* lambdas, size_t n_lambdas, double* out, size_t out_len) {
    if (!h || !out) return fb_set_err(FB200_ERR_BAD_HANDLE, "null handle");
    if (out_len < h->n * n_lambdas) return fb_set_err(FB200_ERR_INVALID_ARGUMENT, "output buffer too small");
    if (n_lambdas == 0) return FB200_OK;
    DeviceGuard guard;
    return gather<double>(h, out, n_lambdas, [&](fb200_ensemble* e, double* p, size_t len) { return fb200_ensemble_flux(e, region, row, lambdas, n_lambdas, p, len); });
}

// ---- one-shot sweep: argument structs in, light curves out ----
int fb200_sweep_plan(const fb200_args_t* args, size_t n_models, const fb200_config_t* cfg, size_t* n_rows, size_t* n_cols) {
    int r = 0, c = 0;
    size_t bad = 0;
    const int rc = plan_ensemble(args, n_models, cfg, &r, &c, nullptr, nullptr, &bad);
    if (rc != FB200_OK) return rc;
    if (n_rows) *n_rows = size_t(r);
    if (n_cols) *n_cols = size_t(c);
    return FB200_OK;
}

int fb200_sweep_run(const fb200_args_t* args, size_t n_models, const fb200_config_t* cfg, const int32_t* devices, int32_t n_devices,
                    double* rows, size_t rows_len, int32_t* status, int32_t* rows_valid, int64_t* picard_iters, double* t_end,
                    fb200_sweep_stats_t* stats, size_t* bad_model) {
    if (bad_model) *bad_model = 0;
    if (!args || !rows || n_models == 0) return fb_set_err(FB200_ERR_INVALID_ARGUMENT, "null argument or empty ensemble");
    int n_rows = 0, n_cols = 0;
    int rc = plan_ensemble(args, n_models, cfg, &n_rows, &n_cols, nullptr, nullptr, bad_model);
    if (rc != FB200_OK) return rc;
    if (rows_len < n_models * size_t(n_rows) * size_t(n_cols)) return fb_set_err(FB200_ERR_INVALID_ARGUMENT, "output buffer too small");
    std::vector<int> devs;
    rc = resolve_devices(devices, n_devices, devs);
    if (rc != FB200_OK) return rc;
    SweepIO io;
    io.rows = rows;
    io.status = status;
    io.rows_valid = rows_valid;
    io.picard_iters = picard_iters;
    io.t_end = t_end;
    rc = build_ensembles(args, n_models, cfg, devs.data(), int(devs.size()), /*keep_initial=*/false, &io, nullptr, bad_model);
    if (rc != FB200_OK) return rc;
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        stats->total_ms = io.total_ms;
        stats->device_ms_max = io.device_ms_max;
        stats->first_launch_ms = io.first_launch_ms;
        stats->resolve_ms = io.resolve_ms;
        stats->uploads_ms = io.uploads_ms;
        stats->model_steps = io.model_steps;
        stats->n_devices = int32_t(std::min(devs.size(), n_models));
        stats->launches = io.launches;
        stats->n_rows = n_rows;
        stats->n_cols = n_cols;
    }
    return FB200_OK;
}

}  // extern "C"
