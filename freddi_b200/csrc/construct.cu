// Construction pipeline of the ensembles (ensemble_impl.hpp): plan -> one device arena per GPU -> host workers resolve the
// models into pinned staging slots -> chunked cudaMemcpyAsync uploads; for a streamed sweep the evolve kernel is already
// running and picks every group up as its upload watermark arrives, and finished groups are read back meanwhile.
// Replaces, per model, the constructors the reference's driver loop runs (cpp/include/application.hpp:17-24:
// `Freddi freddi(args)`), the loop itself (`for (step...)`, :25-43) being the evolve kernel.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <shared_mutex>
#include <thread>

#include "ensemble_impl.hpp"
#include "star.hpp"

extern "C" {  // the kernel builds (evolve.cu / fields.cu compiled once each per build)
size_t fb200_s128_evolve_smem_bytes();
size_t fb200_s128_evolve_scratch_doubles_per_cta();
int fb200_s128_evolve_ctas_per_sm();
int fb200_s128_evolve_resident_ctas_per_sm();
int fb200_s128_evolve_threads();
cudaError_t fb200_s128_launch_evolve(const void* E, int n_ctas, cudaStream_t stream);
cudaError_t fb200_s128_launch_field(const void* E, int field, long long row0, int n_sel, int pad_nan, double* out, cudaStream_t stream);
cudaError_t fb200_s128_launch_flux(const void* E, int region, long long row0, int n_sel, const double* lambdas_dev, int n_lambdas,
                                    const double* phases_dev, double* out, cudaStream_t stream);
cudaError_t fb200_s128_launch_mdot_wind(const void* E, long long row, double* out, cudaStream_t stream);
size_t fb200_s256_evolve_smem_bytes();
size_t fb200_s256_evolve_scratch_doubles_per_cta();
int fb200_s256_evolve_ctas_per_sm();
int fb200_s256_evolve_resident_ctas_per_sm();
int fb200_s256_evolve_threads();
cudaError_t fb200_s256_launch_evolve(const void* E, int n_ctas, cudaStream_t stream);
cudaError_t fb200_s256_launch_field(const void* E, int field, long long row0, int n_sel, int pad_nan, double* out, cudaStream_t stream);
cudaError_t fb200_s256_launch_flux(const void* E, int region, long long row0, int n_sel, const double* lambdas_dev, int n_lambdas,
                                    const double* phases_dev, double* out, cudaStream_t stream);
cudaError_t fb200_s256_launch_mdot_wind(const void* E, long long row, double* out, cudaStream_t stream);
size_t fb200_s512_evolve_smem_bytes();
size_t fb200_s512_evolve_scratch_doubles_per_cta();
int fb200_s512_evolve_ctas_per_sm();
int fb200_s512_evolve_resident_ctas_per_sm();
int fb200_s512_evolve_threads();
cudaError_t fb200_s512_launch_evolve(const void* E, int n_ctas, cudaStream_t stream);
cudaError_t fb200_s512_launch_field(const void* E, int field, long long row0, int n_sel, int pad_nan, double* out, cudaStream_t stream);
cudaError_t fb200_s512_launch_flux(const void* E, int region, long long row0, int n_sel, const double* lambdas_dev, int n_lambdas,
                                    const double* phases_dev, double* out, cudaStream_t stream);
cudaError_t fb200_s512_launch_mdot_wind(const void* E, long long row, double* out, cudaStream_t stream);
size_t fb200_main_evolve_smem_bytes();
size_t fb200_main_evolve_scratch_doubles_per_cta();
int fb200_main_evolve_ctas_per_sm();
int fb200_main_evolve_resident_ctas_per_sm();
int fb200_main_evolve_threads();
cudaError_t fb200_main_launch_evolve(const void* E, int n_ctas, cudaStream_t stream);
cudaError_t fb200_main_launch_field(const void* E, int field, long long row0, int n_sel, int pad_nan, double* out, cudaStream_t stream);
cudaError_t fb200_main_launch_flux(const void* E, int region, long long row0, int n_sel, const double* lambdas_dev, int n_lambdas,
                                    const double* phases_dev, double* out, cudaStream_t stream);
cudaError_t fb200_main_launch_mdot_wind(const void* E, long long row, double* out, cudaStream_t stream);
size_t fb200_big_evolve_smem_bytes();
size_t fb200_big_evolve_scratch_doubles_per_cta();
int fb200_big_evolve_ctas_per_sm();
int fb200_big_evolve_resident_ctas_per_sm();
int fb200_big_evolve_threads();
cudaError_t fb200_big_launch_evolve(const void* E, int n_ctas, cudaStream_t stream);
cudaError_t fb200_big_launch_field(const void* E, int field, long long row0, int n_sel, int pad_nan, double* out, cudaStream_t stream);
cudaError_t fb200_big_launch_flux(const void* E, int region, long long row0, int n_sel, const double* lambdas_dev, int n_lambdas,
                                    const double* phases_dev, double* out, cudaStream_t stream);
cudaError_t fb200_big_launch_mdot_wind(const void* E, long long row, double* out, cudaStream_t stream);
}

namespace fb200 {

// ------------------------------------------------------------------------------------------------ kernel builds
#define FB_BUILD_ROW(v, nx)                                                                                                          \
    KernelBuild {                                                                                                                    \
        #v, nx, fb200_##v##_evolve_smem_bytes, fb200_##v##_evolve_scratch_doubles_per_cta, fb200_##v##_evolve_ctas_per_sm,           \
            fb200_##v##_evolve_resident_ctas_per_sm, fb200_##v##_evolve_threads, fb200_##v##_launch_evolve, fb200_##v##_launch_field, \
            fb200_##v##_launch_flux, fb200_##v##_launch_mdot_wind                                                                    \
    }
const KernelBuild& kernel_build(int index) {
    static const KernelBuild builds[kKernelBuilds] = {FB_BUILD_ROW(s128, 128), FB_BUILD_ROW(s256, 256), FB_BUILD_ROW(s512, 512),
                                                      FB_BUILD_ROW(main, 1024), FB_BUILD_ROW(big, FB200_NX_LIMIT)};
    return builds[index];
}
int kernel_build_index_for(int Nx) {
    static const bool main_only = [] { const char* c = getenv("FB200_BUILD"); return c && std::string(c) == "main"; }();
    for (int k = main_only ? 3 : 0; k < kKernelBuilds; ++k)
        if (Nx <= kernel_build(k).nx_max) return k;
    return kKernelBuilds - 1;
}

// ------------------------------------------------------------------------------------------------ errors
namespace {
thread_local std::string g_err;
}
int fb_set_err(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
int fb_cuda_fail(cudaError_t e, const char* what) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? FB200_ERR_NO_DEVICE : FB200_ERR_CUDA;
}
const std::string& fb_err_msg() { return g_err; }

// ------------------------------------------------------------------------------------------------ host workers
struct HostPool::Impl {
    std::mutex run_mutex;  // one job at a time
    std::mutex m;
    std::condition_variable cv_job, cv_done;
    std::vector<std::thread> threads;
    const std::function<void(unsigned)>* job = nullptr;
    unsigned job_workers = 0;  // pool threads taking part in the current job (the caller is worker 0)
    unsigned long long gen = 0;
    unsigned running = 0;
};

HostPool& HostPool::instance() {
    static HostPool* pool = new HostPool();  // never destroyed: its threads sleep until the process ends
    return *pool;
}

unsigned HostPool::default_threads() {
    if (const char* c = getenv("FB200_HOST_THREADS")) {
        const int v = atoi(c);
        if (v > 0) return unsigned(v);
    }
    const unsigned n = std::thread::hardware_concurrency();
    return n ? n : 1;
}

void HostPool::run(unsigned n_workers, const std::function<void(unsigned)>& body) {
    if (!impl_) {
        static std::mutex init;
        std::lock_guard<std::mutex> g(init);
        if (!impl_) impl_ = new Impl();
    }
    Impl& I = *impl_;
    if (n_workers <= 1) {
        body(0);
        return;
    }
    std::lock_guard<std::mutex> one(I.run_mutex);
    {
        std::unique_lock<std::mutex> lk(I.m);
        while (I.threads.size() + 1 < n_workers) {
            const unsigned idx = unsigned(I.threads.size());
            const unsigned long long seen = I.gen;
            I.threads.emplace_back([&I, idx, seen] {
                unsigned long long last = seen;
                for (;;) {
                    const std::function<void(unsigned)>* job = nullptr;
                    {
                        std::unique_lock<std::mutex> lk(I.m);
                        I.cv_job.wait(lk, [&] { return I.gen != last; });
                        last = I.gen;
                        if (idx < I.job_workers) job = I.job;
                    }
                    if (job) {
                        (*job)(idx + 1);
                        std::lock_guard<std::mutex> lk(I.m);
                        if (--I.running == 0) I.cv_done.notify_all();
                    }
                }
            });
            I.threads.back().detach();
        }
        I.job = &body;
        I.job_workers = n_workers - 1;
        I.running = n_workers - 1;
        ++I.gen;
    }
    I.cv_job.notify_all();
    body(0);
    std::unique_lock<std::mutex> lk(I.m);
    I.cv_done.wait(lk, [&] { return I.running == 0; });
    I.job = nullptr;
}

// ------------------------------------------------------------------------------------------------ caches
// A streamed sweep launches its kernel before the models are uploaded: from the launch until its last upload is complete the
// kernel depends on this process's host threads.  Driver calls that may wait for the device to go idle — page-locked host
// allocation, growth of a memory pool, the frees of fb200_trim — issued by ANOTHER host thread in that window would keep the
// uploads' driver calls waiting behind them while the kernel waits for the uploads (seen as 10 s work-item time-outs with two
// sweeps at once).  So every streamed sweep holds this lock shared during its upload window (milliseconds), and the allocation
// paths take it exclusively: an allocation waits for the open windows to close, never for a kernel that waits for it.
std::shared_mutex g_upload_window;

namespace {
struct Block {
    void* p;
    size_t bytes;
    bool in_use;
    int device;
};
std::mutex g_cache_mutex;
std::vector<Block> g_pinned;
bool cache_disabled() {
    static const bool off = getenv("FB200_NO_CACHE") != nullptr;
    return off;
}
}  // namespace

void* pinned_acquire(size_t bytes) {
    size_t want = size_t(1) << 20;
    while (want < bytes) want <<= 1;
    {
        std::lock_guard<std::mutex> g(g_cache_mutex);
        for (Block& b : g_pinned)
            if (!b.in_use && b.bytes == want) {
                b.in_use = true;
                return b.p;
            }
    }
    void* p = nullptr;
    {
        std::unique_lock<std::shared_mutex> no_open_window(g_upload_window);
        if (cudaHostAlloc(&p, want, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
    }
    std::lock_guard<std::mutex> g(g_cache_mutex);
    g_pinned.push_back(Block{p, want, true, -1});
    return p;
}

void pinned_release(void* block) {
    if (!block) return;
    std::lock_guard<std::mutex> g(g_cache_mutex);
    for (size_t i = 0; i < g_pinned.size(); ++i)
        if (g_pinned[i].p == block) {
            if (cache_disabled()) {
                cudaFreeHost(block);
                g_pinned.erase(g_pinned.begin() + i);
            } else {
                g_pinned[i].in_use = false;
            }
            return;
        }
}

// Device arenas come from a stream-ordered memory pool of the library's own per device (cudaMallocFromPoolAsync / cudaFreeAsync):
// neither call synchronises the device — cudaFree does, and with a second ensemble's kernel waiting for ITS uploads at that
// moment (two sweeps at once in one process) the uploads queue up behind the blocked free until the kernel's waits time out.
// The pool keeps what is freed (release threshold: unlimited) until fb200_trim(); FB200_NO_CACHE=1: nothing is kept.
namespace {
std::mutex g_pool_mutex;
cudaMemPool_t g_pools[64] = {};
int device_pool(int device, cudaMemPool_t* out) {
    std::lock_guard<std::mutex> g(g_pool_mutex);
    cudaMemPool_t& pool = g_pools[device & 63];
    if (!pool) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        FB_CUDA(cudaMemPoolCreate(&pool, &props));
        unsigned long long keep = cache_disabled() ? 0ull : ~0ull;
        FB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    *out = pool;
    return FB200_OK;
}
}  // namespace

// the calling thread has `device` selected; `stream` is an idle stream of that device
int arena_acquire(int device, size_t bytes, cudaStream_t stream, void** p) {
    cudaMemPool_t pool;
    int rc = device_pool(device, &pool);
    if (rc != FB200_OK) return rc;
    std::unique_lock<std::shared_mutex> no_open_window(g_upload_window);  // (the pool may have to grow)
    cudaError_t ce = cudaMallocFromPoolAsync(p, bytes, pool, stream);
    if (ce == cudaErrorMemoryAllocation) {  // give back what the pool holds idle and try once more
        cudaGetLastError();
        cudaMemPoolTrimTo(pool, 0);
        ce = cudaMallocFromPoolAsync(p, bytes, pool, stream);
    }
    if (ce != cudaSuccess) return fb_cuda_fail(ce, "cudaMallocFromPoolAsync (ensemble arena)");
    // the block is used from three streams: let the allocation complete (the stream is idle) instead of ordering each of them behind it
    FB_CUDA(cudaStreamSynchronize(stream));
    return FB200_OK;
}

void arena_release(void* p, cudaStream_t stream) {
    if (p) cudaFreeAsync(p, stream);
}

void caches_trim() {
    std::unique_lock<std::shared_mutex> no_open_window(g_upload_window);
    {
        std::lock_guard<std::mutex> g(g_pool_mutex);
        for (cudaMemPool_t pool : g_pools)
            if (pool) cudaMemPoolTrimTo(pool, 0);
    }
    std::lock_guard<std::mutex> g(g_cache_mutex);
    for (size_t i = 0; i < g_pinned.size();)
        if (!g_pinned[i].in_use) {
            cudaFreeHost(g_pinned[i].p);
            g_pinned.erase(g_pinned.begin() + i);
        } else {
            ++i;
        }
}

// ------------------------------------------------------------------------------------------------ plan
namespace {

struct Plan {
    int Nx = 0, n_rows = 0, n_cols = 0, max_Nt = 0;
    int build = 3;  // kernel_build index
    bool is_ns = false, big = false, has_A = false, has_B = false, has_C = false, has_dyn = false;
};

int n_cols_of(const fb200_config_t& cfg, bool is_ns) {
    return FB200_N_BH_COLS + (cfg.n_lambdas + cfg.n_passbands) * ((cfg.cold_disk ? 2 : 1) + (cfg.star_flux == 1 ? 3 : 0)) +
           (is_ns ? FB200_N_NS_COLS : 0);
}

int make_plan(const fb200_args_t* args, size_t n, const fb200_config_t& cfg, Plan& P, size_t* bad_model) {
    const ModelPlan p0 = plan_of(args[0]);
    P.Nx = p0.Nx;
    P.is_ns = p0.is_ns;
    for (size_t i = 0; i < n; ++i) {
        const ModelPlan p = plan_of(args[i]);
        if (p.Nx != P.Nx || p.is_ns != P.is_ns) {
            if (bad_model) *bad_model = i;
            return fb_set_err(FB200_ERR_INVALID_ARGUMENT, "all models of one ensemble must share Nx and the model class (BH / NS)");
        }
        P.max_Nt = std::max(P.max_Nt, p.Nt);
        P.has_A |= p.has_A;
        P.has_B |= p.has_B;
        P.has_C |= p.has_C;
        P.has_dyn |= p.dynamic;
    }
    P.n_rows = P.max_Nt + 1;
    P.n_cols = n_cols_of(cfg, P.is_ns);
    P.big = P.Nx > 1024;
    P.build = kernel_build_index_for(P.Nx);
    return FB200_OK;
}

// scheduling groups / upload chunks: boundaries, every group but the last a multiple of 16 models (EnsembleDev::ready)
std::vector<int> partition_groups(size_t n, int first, int size) {
    std::vector<int> b{0};
    if (n > size_t(first) + size / 2) {
        b.push_back(first);
        while (n - size_t(b.back()) > size_t(size) + size / 2) b.push_back(b.back() + size);
    }
    b.push_back(int(n));
    return b;
}

struct SlotLayout {  // one staging slot: the arrays of up to `cap` models, back to back, 256-byte aligned
    size_t cap = 0, o_models = 0, o_state = 0, o_h = 0, o_F = 0, o_A = 0, o_B = 0, o_C = 0, o_Fm = 0, o_dFm = 0, o_d2Fm = 0, bytes = 0;
};

SlotLayout slot_layout(const Plan& P, size_t cap) {
    SlotLayout L;
    L.cap = cap;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        off = (off + 255) & ~size_t(255);
        const size_t o = off;
        off += bytes;
        return o;
    };
    const size_t arr = cap * size_t(P.Nx) * sizeof(double);
    L.o_models = take(cap * sizeof(fb200_model_dev_t));
    L.o_state = take(cap * sizeof(ModelState));
    L.o_h = take(arr);
    L.o_F = take(arr);
    if (P.has_A) L.o_A = take(arr);
    if (P.has_B) L.o_B = take(arr);
    if (P.has_C) L.o_C = take(arr);
    if (P.is_ns) {
        L.o_Fm = take(arr);
        L.o_dFm = take(arr);
        L.o_d2Fm = take(arr);
    }
    L.bytes = off;
    return L;
}

constexpr int kSlots = 3;

struct DeviceBuild {
    fb200_ensemble* e = nullptr;
    int device = 0;
    ArgView args;
    size_t global0 = 0, gstride = 1;  // local model j = model global0 + j * gstride of the caller
    std::vector<int> chunk_begin;     // [n_chunks + 1]
    std::unique_ptr<std::atomic<int>[]> chunk_left;
    std::atomic<int> uploaded{0};     // chunks whose upload is complete (their staging slot is free again)
    SlotLayout L;
    char* slots[kSlots] = {nullptr, nullptr, nullptr};
    cudaEvent_t slot_ev[kSlots] = {nullptr, nullptr, nullptr};
    int enqueued = 0, polled = 0;
    int* marks = nullptr;             // pinned: cumulative model counts per chunk (+ a -1 at the end): sources of the watermark copies
    void* marks_block = nullptr;
    int next_block = 0;               // streamed sweep: read-back blocks whose rows have been sent to the host
    int rc = FB200_OK;
    std::string err;
    double t_launch_ms = 0., t_uploaded_ms = 0., device_ms = 0.;
};

struct Shared {
    Plan P;
    fb200_config_t cfg{};
    std::vector<PassbandTable> passbands;
    std::vector<double> star_table;             // all star geometries back to back
    std::vector<long long> star_off;            // per caller model
    std::vector<int> star_ntri;
    int star_ntri_max = 0;
    const double* lx_table = nullptr;           // host blob of the band-integral table (lx_table.h), or nullptr: every ring walks
    size_t lx_doubles = 0;
    double lx_r = 0.;                           // its band ratio emax / emin
    bool keep_initial = false;
    SweepIO* io = nullptr;
    size_t n_total = 0;
    int n_dev = 0;
    std::atomic<bool> abort{false};
    std::atomic<int> allocated{0};  // devices whose arena is in place (no upload window opens before all of them are: g_upload_window)
    std::mutex err_mutex;
    int err_code = FB200_OK;
    size_t err_model = ~size_t(0);
    std::string err_msg;
    std::chrono::steady_clock::time_point t0;
    double ms() const { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
    void fail(int code, const std::string& msg, size_t model) {
        std::lock_guard<std::mutex> g(err_mutex);
        if (err_code == FB200_OK || model < err_model) {
            err_code = code;
            err_msg = msg;
            err_model = model;
        }
        abort.store(true);
    }
};

struct DevCaps {
    bool known = false;
    int sm_count = 0, smem_optin = 0;
    int resident[kKernelBuilds] = {-2, -2, -2, -2, -2};  // CTAs per SM the occupancy calculator grants each kernel build (-2: not asked yet)
};
std::mutex g_caps_mutex;
DevCaps g_caps[64];

int device_caps(int device, int build, DevCaps& out) {  // the calling thread has `device` selected
    std::lock_guard<std::mutex> g(g_caps_mutex);
    DevCaps& c = g_caps[device & 63];
    if (!c.known) {
        FB_CUDA(cudaDeviceGetAttribute(&c.sm_count, cudaDevAttrMultiProcessorCount, device));
        FB_CUDA(cudaDeviceGetAttribute(&c.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        c.known = true;
    }
    const KernelBuild& kb = kernel_build(build);
    if (c.resident[build] == -2) c.resident[build] = (size_t(c.smem_optin) >= kb.smem_bytes()) ? kb.resident_ctas_per_sm() : 0;
    out = c;
    return FB200_OK;
}

// Carves the arena (base == nullptr: only measures).  Returns the bytes needed.
size_t carve(fb200_ensemble* e, const Shared& sh, char* base, size_t scratch_per_cta) {
    const Plan& P = sh.P;
    EnsembleDev& D = e->dev;
    size_t off = 0;
    auto take = [&](size_t bytes) -> void* {
        off = (off + 255) & ~size_t(255);
        void* p = base ? base + off : nullptr;
        off += bytes;
        return p;
    };
    const size_t n = e->n, NN = n * size_t(P.Nx) * sizeof(double);
    // the streamed arrays first (uploads of different chunks never share a 128-byte line: chunk sizes are multiples of 16 models)
    auto* models = static_cast<fb200_model_dev_t*>(take(n * sizeof(fb200_model_dev_t)));
    auto* state = static_cast<ModelState*>(take(n * sizeof(ModelState)));
    auto* h = static_cast<double*>(take(NN));
    auto* F = static_cast<double*>(take(NN));
    auto* wA = P.has_A ? static_cast<double*>(take(NN)) : nullptr;
    auto* wB = P.has_B ? static_cast<double*>(take(NN)) : nullptr;
    auto* wC = P.has_C ? static_cast<double*>(take(NN)) : nullptr;
    auto* Fm = P.is_ns ? static_cast<double*>(take(NN)) : nullptr;
    auto* dFm = P.is_ns ? static_cast<double*>(take(NN)) : nullptr;
    auto* d2Fm = P.is_ns ? static_cast<double*>(take(NN)) : nullptr;
    auto* K = static_cast<double*>(take(NN));
    auto* rows = static_cast<double*>(take(n * size_t(P.n_rows) * P.n_cols * sizeof(double)));
    auto* queue = static_cast<int*>(take(64));
    auto* progress = static_cast<int*>(take(n * sizeof(int)));
    auto* failed = static_cast<int*>(take(n * sizeof(int)));
    auto* wind_rec = P.has_dyn ? static_cast<WindRecord*>(take(n * size_t(P.n_rows) * sizeof(WindRecord))) : nullptr;
    auto* cta_scratch = static_cast<double*>(take(size_t(D.max_ctas) * scratch_per_cta * sizeof(double)));
    auto* star_T = sh.star_ntri_max > 0 ? static_cast<double*>(take(size_t(D.max_ctas) * sh.star_ntri_max * sizeof(double))) : nullptr;
    auto* prof = static_cast<long long*>(take(16 * sizeof(long long)));
    double *pb_a = nullptr, *pb_w = nullptr;
    if (sh.cfg.n_passbands > 0) {
        size_t len = 0;
        for (const PassbandTable& t : sh.passbands) len += t.a.size();
        pb_a = static_cast<double*>(take(len * sizeof(double)));
        pb_w = static_cast<double*>(take(len * sizeof(double)));
    }
    auto* star_data = sh.star_table.empty() ? nullptr : static_cast<double*>(take(sh.star_table.size() * sizeof(double)));
    auto* lx_table = sh.lx_table ? static_cast<double*>(take(sh.lx_doubles * sizeof(double))) : nullptr;
    double* Fsnap = nullptr;
    int* bounds = nullptr;
    if (sh.cfg.record_F) {
        Fsnap = static_cast<double*>(take(n * size_t(P.n_rows) * P.Nx * sizeof(double)));
        bounds = static_cast<int*>(take(n * size_t(P.n_rows) * 2 * sizeof(int)));
    }
    double* F0 = nullptr;
    ModelState* state0 = nullptr;
    if (sh.keep_initial) {
        F0 = static_cast<double*>(take(NN));
        state0 = static_cast<ModelState*>(take(n * sizeof(ModelState)));
    }
    int *ready = nullptr, *block_count = nullptr;
    const size_t n_blocks = (n + kDoneBlock - 1) / kDoneBlock;
    if (sh.io) {
        ready = static_cast<int*>(take(64));
        block_count = static_cast<int*>(take(n_blocks * sizeof(int)));
    }
    if (base) {
        D.models = models; D.state = state; D.h = h; D.F = F; D.K = K;
        D.wA = wA; D.wB = wB; D.wCs = wC; D.Fmagn = Fm; D.dFmagn_dh = dFm; D.d2Fmagn_dh2 = d2Fm;
        D.rows = rows; D.queue = queue; D.progress = progress; D.failed = failed; D.wind_rec = wind_rec;
        D.cta_scratch = cta_scratch; D.star_T = star_T; D.prof = prof;
        D.cfg.pb_a = pb_a; D.cfg.pb_w = pb_w; D.cfg.star_data = star_data; D.cfg.lx_table = lx_table;
        D.Fsnap = Fsnap; D.bounds = bounds;
        e->F0_dev = F0; e->state0_dev = state0;
        e->ready_dev = ready;
        D.ready = ready;
        D.block_count = block_count;
    }
    return off + 256;
}

#define FB_TRYB(call)                                             \
    do {                                                          \
        cudaError_t _e = (call);                                  \
        if (_e != cudaSuccess) {                                  \
            B.rc = fb_cuda_fail(_e, #call);                       \
            B.err = fb_err_msg();                                 \
            return B.rc;                                          \
        }                                                         \
    } while (0)

// streams, arena, tables and the clearing memsets of one device's ensemble
int device_alloc(DeviceBuild& B, Shared& sh) {
    const Plan& P = sh.P;
    auto fail = [&](int code, const std::string& msg) { B.rc = code; B.err = msg; return code; };
    int ndev = 0;
    {
        cudaError_t ce = cudaGetDeviceCount(&ndev);
        if (ce != cudaSuccess || ndev == 0) {
            std::string why = (ce != cudaSuccess) ? std::string("cudaGetDeviceCount: ") + cudaGetErrorString(ce) : std::string("no CUDA device");
            cudaGetLastError();
            return fail(FB200_ERR_NO_DEVICE, "no CUDA device: " + why + " (there is no CPU fallback)");
        }
    }
    if (B.device < 0 || B.device >= ndev) return fail(FB200_ERR_INVALID_ARGUMENT, "device ordinal out of range");
    fb200_ensemble* e = B.e;
    FB_TRYB(cudaSetDevice(B.device));
    DevCaps caps;
    const KernelBuild& kb = kernel_build(P.build);
    e->kb = &kb;
    if (device_caps(B.device, P.build, caps) != FB200_OK) return fail(FB200_ERR_CUDA, fb_err_msg());
    e->sm_count = caps.sm_count;
    if (size_t(caps.smem_optin) < kb.smem_bytes())
        return fail(FB200_ERR_UNSUPPORTED, "device offers too little shared memory per block for the sm_100a kernels");
    FB_TRYB(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    FB_TRYB(cudaStreamCreateWithFlags(&e->copy_in, cudaStreamNonBlocking));
    FB_TRYB(cudaStreamCreateWithFlags(&e->copy_out, cudaStreamNonBlocking));
    FB_TRYB(cudaEventCreate(&e->ev0));
    FB_TRYB(cudaEventCreate(&e->ev1));
    for (int s = 0; s < kSlots; ++s) FB_TRYB(cudaEventCreateWithFlags(&B.slot_ev[s], cudaEventDisableTiming));

    EnsembleDev& D = e->dev;
    const fb200_config_t& cfg = sh.cfg;
    D.n_models = int(e->n);
    D.Nx = P.Nx;
    D.cfg.n_lambdas = cfg.n_lambdas;
    D.cfg.cold_disk = cfg.cold_disk;
    D.cfg.compute_lx = cfg.compute_lx;
    D.cfg.record_F = cfg.record_F;
    D.cfg.lx_prune = cfg.exact_lx_walk ? 0 : 1;
    D.cfg.n_cols = e->n_cols;
    D.cfg.n_rows = e->n_rows;
    for (int k = 0; k < FB200_MAX_LAMBDAS; ++k) D.cfg.lambdas[k] = (k < cfg.n_lambdas) ? cfg.lambdas[k] : 0.;
    D.cfg.n_passbands = cfg.n_passbands;
    for (int k = 0; k < FB200_MAX_PASSBANDS + 2; ++k) D.cfg.pb_off[k] = 0;
    for (int k = 0; k < FB200_MAX_PASSBANDS; ++k) D.cfg.pb_tdnu[k] = 1.;
    D.cfg.star_flux = (cfg.star_flux == 1) ? 1 : 0;  // 2: geometry for fb200_ensemble_flux_rows(region 2) only, no columns
    D.star_ntri_max = sh.star_ntri_max;
    {
        // the chunked scheduler lets a CTA wait for another one: never launch more CTAs than can be resident
        int resident = caps.resident[P.build];
        if (resident < 1) return fail(FB200_ERR_UNSUPPORTED, "the evolve kernel does not fit on this device");
        const int designed = kb.ctas_per_sm();
        if (resident > designed) resident = designed;
        if (P.big && size_t(e->sm_count) * resident > e->n)  // 3 MB of scratch per CTA: no more CTAs than models
            resident = int((e->n + e->sm_count - 1) / e->sm_count);
        D.max_ctas = e->sm_count * resident;
    }
    {
        long long ms = cfg.wait_timeout_ms > 0 ? cfg.wait_timeout_ms : 10000;
        if (const char* c = getenv("FB200_WAIT_TIMEOUT_US")) D.wait_timeout_ns = 1000ll * atoll(c);  // test hook: forces the give-up path
        else D.wait_timeout_ns = ms * 1000000ll;
    }
    D.block_done = nullptr;

    const size_t scratch_per_cta = kb.scratch_doubles_per_cta();
    const size_t need = carve(e, sh, nullptr, scratch_per_cta);
    void* base = nullptr;
    if (arena_acquire(B.device, need, e->stream, &base) != FB200_OK) return fail(FB200_ERR_CUDA, fb_err_msg());
    e->arena = base;
    e->arena_bytes = need;
    carve(e, sh, static_cast<char*>(base), scratch_per_cta);

    if (cfg.n_passbands > 0) {  // all tables back to back: a_j, then w_j
        std::vector<double> ha, hw;
        for (int p = 0; p < cfg.n_passbands; ++p) {
            D.cfg.pb_off[p] = int(ha.size());
            ha.insert(ha.end(), sh.passbands[p].a.begin(), sh.passbands[p].a.end());
            hw.insert(hw.end(), sh.passbands[p].w.begin(), sh.passbands[p].w.end());
            D.cfg.pb_tdnu[p] = sh.passbands[p].t_dnu;
        }
        for (int p = cfg.n_passbands; p < FB200_MAX_PASSBANDS + 2; ++p) D.cfg.pb_off[p] = int(ha.size());
        FB_TRYB(cudaMemcpy(const_cast<double*>(D.cfg.pb_a), ha.data(), sizeof(double) * ha.size(), cudaMemcpyHostToDevice));
        FB_TRYB(cudaMemcpy(const_cast<double*>(D.cfg.pb_w), hw.data(), sizeof(double) * hw.size(), cudaMemcpyHostToDevice));
    }
    if (!sh.star_table.empty())
        FB_TRYB(cudaMemcpy(const_cast<double*>(D.cfg.star_data), sh.star_table.data(), sizeof(double) * sh.star_table.size(), cudaMemcpyHostToDevice));
    if (sh.lx_table)  // 40-70 KB; the host blob lives as long as the process
        FB_TRYB(cudaMemcpyAsync(const_cast<double*>(D.cfg.lx_table), sh.lx_table, sizeof(double) * sh.lx_doubles, cudaMemcpyHostToDevice, e->stream));

    const size_t n = e->n;
    FB_TRYB(cudaMemsetAsync(D.failed, 0, sizeof(int) * n, e->stream));
    FB_TRYB(cudaMemsetAsync(D.prof, 0, 16 * sizeof(long long), e->stream));
    if (D.wind_rec) FB_TRYB(cudaMemsetAsync(D.wind_rec, 0, n * size_t(e->n_rows) * sizeof(WindRecord), e->stream));
    FB_TRYB(cudaMemsetAsync(D.rows, 0xFF, n * size_t(e->n_rows) * e->n_cols * sizeof(double), e->stream));  // all-ones = NaN: rows never written stay NaN
    if (D.Fsnap) {
        FB_TRYB(cudaMemsetAsync(D.Fsnap, 0xFF, n * size_t(e->n_rows) * P.Nx * sizeof(double), e->stream));
        FB_TRYB(cudaMemsetAsync(D.bounds, 0xFF, n * size_t(e->n_rows) * 2 * sizeof(int), e->stream));
    }
    if (sh.io) {
        // the watermark is cleared on the stream that will raise it; the launch waits for that (the arena is recycled memory)
        FB_TRYB(cudaMemsetAsync(e->ready_dev, 0, sizeof(int), e->copy_in));
        FB_TRYB(cudaEventRecord(B.slot_ev[0], e->copy_in));
        FB_TRYB(cudaStreamWaitEvent(e->stream, B.slot_ev[0], 0));
        const size_t n_blocks = (n + kDoneBlock - 1) / kDoneBlock;
        FB_TRYB(cudaMemsetAsync(D.block_count, 0, sizeof(int) * n_blocks, e->stream));
        e->pinned_block = pinned_acquire(sizeof(int) * n_blocks);
        if (!e->pinned_block) return fail(FB200_ERR_CUDA, "pinned host memory for the group flags could not be allocated");
        e->block_done_host = static_cast<int*>(e->pinned_block);
        std::memset(e->block_done_host, 0, sizeof(int) * n_blocks);
        void* dptr = nullptr;
        FB_TRYB(cudaHostGetDevicePointer(&dptr, e->block_done_host, 0));
        D.block_done = static_cast<int*>(dptr);
    }
    return FB200_OK;
}

// resolved model -> its place in the staging slot of its chunk
void pack_model(DeviceBuild& B, const Shared& sh, int chunk, size_t k /*index inside the chunk*/, Resolved& r, size_t global) {
    const Plan& P = sh.P;
    const SlotLayout& L = B.L;
    char* slot = B.slots[chunk % kSlots];
    const size_t Nx = size_t(P.Nx), arr = Nx * sizeof(double);
    if (!sh.star_off.empty()) {
        r.dev.star_off = sh.star_off[global];
        r.dev.star_ntri = sh.star_ntri[global];
    }
    r.dev.lx_tab = (sh.lx_table && fabs(r.dev.emax / r.dev.emin - sh.lx_r) <= 4e-16 * sh.lx_r) ? 1 : 0;
    reinterpret_cast<fb200_model_dev_t*>(slot + L.o_models)[k] = r.dev;
    ModelState& s = reinterpret_cast<ModelState*>(slot + L.o_state)[k];
    std::memset(&s, 0, sizeof(s));
    s.t = r.dev.t0;
    s.F_in = r.dev.F_in0;
    s.Mdot_in_prev = -INFINITY;  // cpp/include/freddi_state.hpp:245
    s.first = r.dev.first0;
    s.last = P.Nx - 1;
    std::memcpy(slot + L.o_h + k * arr, r.h.data(), arr);
    std::memcpy(slot + L.o_F + k * arr, r.F.data(), arr);
    auto put = [&](size_t off, const std::vector<double>& v) {
        if (v.empty()) std::memset(slot + off + k * arr, 0, arr);
        else std::memcpy(slot + off + k * arr, v.data(), arr);
    };
    if (P.has_A) put(L.o_A, r.wA);
    if (P.has_B) put(L.o_B, r.wB);
    if (P.has_C) {  // static wind C + d2Fmagn_dh2 (FreddiNeutronStarEvolution::windC, cpp/src/ns/ns_evolution.cpp:487-493)
        double* c = reinterpret_cast<double*>(slot + L.o_C + k * arr);
        if (r.wC.empty()) std::memset(c, 0, arr);
        else std::memcpy(c, r.wC.data(), arr);
        if (!r.d2Fmagn_dh2.empty())
            for (size_t i = 0; i < Nx; ++i) c[i] += r.d2Fmagn_dh2[i];
    }
    if (P.is_ns) {
        put(L.o_Fm, r.Fmagn);
        put(L.o_dFm, r.dFmagn_dh);
        put(L.o_d2Fm, r.d2Fmagn_dh2);
    }
}

int enqueue_upload(DeviceBuild& B, const Shared& sh, int c) {
    fb200_ensemble* e = B.e;
    const EnsembleDev& D = e->dev;
    const Plan& P = sh.P;
    const SlotLayout& L = B.L;
    const char* slot = B.slots[c % kSlots];
    const size_t m0 = size_t(B.chunk_begin[c]), cnt = size_t(B.chunk_begin[c + 1]) - m0;
    const size_t Nx = size_t(P.Nx), arr = cnt * Nx * sizeof(double), o = m0 * Nx;
    cudaStream_t s = e->copy_in;
    FB_TRYB(cudaMemcpyAsync(const_cast<fb200_model_dev_t*>(D.models) + m0, slot + L.o_models, cnt * sizeof(fb200_model_dev_t), cudaMemcpyHostToDevice, s));
    FB_TRYB(cudaMemcpyAsync(D.state + m0, slot + L.o_state, cnt * sizeof(ModelState), cudaMemcpyHostToDevice, s));
    FB_TRYB(cudaMemcpyAsync(const_cast<double*>(D.h) + o, slot + L.o_h, arr, cudaMemcpyHostToDevice, s));
    FB_TRYB(cudaMemcpyAsync(D.F + o, slot + L.o_F, arr, cudaMemcpyHostToDevice, s));
    if (P.has_A) FB_TRYB(cudaMemcpyAsync(const_cast<double*>(D.wA) + o, slot + L.o_A, arr, cudaMemcpyHostToDevice, s));
    if (P.has_B) FB_TRYB(cudaMemcpyAsync(const_cast<double*>(D.wB) + o, slot + L.o_B, arr, cudaMemcpyHostToDevice, s));
    if (P.has_C) FB_TRYB(cudaMemcpyAsync(const_cast<double*>(D.wCs) + o, slot + L.o_C, arr, cudaMemcpyHostToDevice, s));
    if (P.is_ns) {
        FB_TRYB(cudaMemcpyAsync(const_cast<double*>(D.Fmagn) + o, slot + L.o_Fm, arr, cudaMemcpyHostToDevice, s));
        FB_TRYB(cudaMemcpyAsync(const_cast<double*>(D.dFmagn_dh) + o, slot + L.o_dFm, arr, cudaMemcpyHostToDevice, s));
        FB_TRYB(cudaMemcpyAsync(const_cast<double*>(D.d2Fmagn_dh2) + o, slot + L.o_d2Fm, arr, cudaMemcpyHostToDevice, s));
    }
    if (e->F0_dev) {
        FB_TRYB(cudaMemcpyAsync(e->F0_dev + o, D.F + o, arr, cudaMemcpyDeviceToDevice, s));
        FB_TRYB(cudaMemcpyAsync(e->state0_dev + m0, D.state + m0, cnt * sizeof(ModelState), cudaMemcpyDeviceToDevice, s));
    }
    if (sh.io)  // the watermark, behind the arrays it announces
        FB_TRYB(cudaMemcpyAsync(e->ready_dev, B.marks + c, sizeof(int), cudaMemcpyHostToDevice, s));
    FB_TRYB(cudaEventRecord(B.slot_ev[c % kSlots], s));
    ++B.enqueued;
    return FB200_OK;
}

void poll_uploads(DeviceBuild& B) {
    while (B.polled < B.enqueued && cudaEventQuery(B.slot_ev[B.polled % kSlots]) == cudaSuccess) {
        ++B.polled;
        B.uploaded.store(B.polled, std::memory_order_release);
    }
}

// rows of the read-back blocks [B.next_block, upto) -> the caller's buffer, one copy for the whole run of blocks (models of this
// device are every gstride-th of the caller's)
int send_blocks(DeviceBuild& B, const Shared& sh, int upto) {
    fb200_ensemble* e = B.e;
    if (upto <= B.next_block) return FB200_OK;
    const size_t row_bytes = size_t(e->n_rows) * e->n_cols * sizeof(double);
    const size_t m0 = size_t(B.next_block) * kDoneBlock, m1 = std::min(size_t(upto) * kDoneBlock, e->n), cnt = m1 - m0;
    char* dst = reinterpret_cast<char*>(sh.io->rows) + (B.global0 + m0 * B.gstride) * row_bytes;
    const char* src = reinterpret_cast<const char*>(e->dev.rows) + m0 * row_bytes;
    if (B.gstride == 1) FB_TRYB(cudaMemcpyAsync(dst, src, cnt * row_bytes, cudaMemcpyDeviceToHost, e->copy_out));
    else FB_TRYB(cudaMemcpy2DAsync(dst, B.gstride * row_bytes, src, row_bytes, row_bytes, cnt, cudaMemcpyDeviceToHost, e->copy_out));
    B.next_block = upto;
    return FB200_OK;
}

int poll_blocks(DeviceBuild& B, const Shared& sh) {
    const volatile int* done = B.e->block_done_host;
    int upto = B.next_block;
    const int n_blocks = int((B.e->n + kDoneBlock - 1) / kDoneBlock);
    while (upto < n_blocks && done[upto]) ++upto;
    if (upto > B.next_block) {
        std::atomic_thread_fence(std::memory_order_acquire);
        return send_blocks(B, sh, upto);
    }
    return FB200_OK;
}

void nap() { std::this_thread::sleep_for(std::chrono::microseconds(20)); }

// the driver thread of one device
int drive_device(DeviceBuild& B, Shared& sh) {
    const int alloc_rc = device_alloc(B, sh);
    sh.allocated.fetch_add(1);
    if (alloc_rc != FB200_OK) {
        sh.fail(B.rc, B.err, ~size_t(0));
        return B.rc;
    }
    // every device of this call allocates before any of them opens its upload window (an allocation waits for open windows)
    while (sh.allocated.load() < sh.n_dev && !sh.abort.load()) nap();
    fb200_ensemble* e = B.e;
    const int n_chunks = int(B.chunk_begin.size()) - 1;
    auto cuda_rc = [&](int rc) {
        if (rc != FB200_OK) sh.fail(B.rc, B.err, ~size_t(0));
        return rc;
    };
    std::shared_lock<std::shared_mutex> upload_window(g_upload_window, std::defer_lock);
    if (sh.io) {
        upload_window.lock();  // from the launch until the last upload is complete (see g_upload_window)
        if (ensemble_launch(e, -1) != FB200_OK) {
            B.rc = FB200_ERR_CUDA;
            B.err = fb_err_msg();
            sh.fail(B.rc, B.err, ~size_t(0));
            return B.rc;
        }
        B.t_launch_ms = sh.ms();
    }
    bool aborted = false;
    for (int c = 0; c < n_chunks && !aborted; ++c) {
        while (B.chunk_left[c].load(std::memory_order_acquire) != 0) {
            if (sh.abort.load()) { aborted = true; break; }
            poll_uploads(B);
            if (sh.io && cuda_rc(poll_blocks(B, sh)) != FB200_OK) { aborted = true; break; }
            nap();
        }
        if (aborted) break;
        if (cuda_rc(enqueue_upload(B, sh, c)) != FB200_OK) { aborted = true; break; }
        poll_uploads(B);
    }
    if (aborted || sh.abort.load()) {
        if (sh.io) {  // withdraw the launch: the kernel's work items stop waiting for uploads
            cudaMemcpyAsync(e->ready_dev, B.marks + n_chunks, sizeof(int), cudaMemcpyHostToDevice, e->copy_in);
        }
        cudaStreamSynchronize(e->copy_in);
        cudaStreamSynchronize(e->stream);
        B.uploaded.store(n_chunks);
        return FB200_ERR_RUNTIME;
    }
    FB_TRYB(cudaStreamSynchronize(e->copy_in));
    if (upload_window.owns_lock()) upload_window.unlock();  // the kernel no longer depends on this process's host threads
    B.polled = B.enqueued;
    B.uploaded.store(n_chunks, std::memory_order_release);
    B.t_uploaded_ms = sh.ms();
    if (!sh.io) {
        FB_TRYB(cudaStreamSynchronize(e->stream));  // the clearing memsets
        return FB200_OK;
    }
    // streamed sweep: read finished groups back while the rest still runs
    const bool dbg = getenv("FB200_DEBUG") != nullptr;
    for (;;) {
        const cudaError_t q = cudaEventQuery(e->ev1);
        if (q == cudaSuccess) break;
        if (q != cudaErrorNotReady) {
            B.rc = fb_cuda_fail(q, "fb200_evolve_kernel");
            B.err = fb_err_msg();
            sh.fail(B.rc, B.err, ~size_t(0));
            return B.rc;
        }
        if (cuda_rc(poll_blocks(B, sh)) != FB200_OK) return B.rc;
        nap();
    }
    if (dbg) fprintf(stderr, "[fb200] device %d: kernel done at %8.2f ms, %d of %d read-back blocks already sent\n", B.device, sh.ms(), B.next_block,
                     int((e->n + kDoneBlock - 1) / kDoneBlock));
    if (cuda_rc(send_blocks(B, sh, int((e->n + kDoneBlock - 1) / kDoneBlock))) != FB200_OK) return B.rc;
    {
        float ms = 0.f;
        FB_TRYB(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
        B.device_ms = ms;
    }
    std::vector<ModelState> hs(e->n);
    std::vector<int> failed(e->n);
    FB_TRYB(cudaMemcpyAsync(hs.data(), e->dev.state, sizeof(ModelState) * e->n, cudaMemcpyDeviceToHost, e->stream));
    FB_TRYB(cudaMemcpyAsync(failed.data(), e->dev.failed, sizeof(int) * e->n, cudaMemcpyDeviceToHost, e->stream));
    FB_TRYB(cudaStreamSynchronize(e->stream));
    long long steps = 0;
    for (size_t j = 0; j < e->n; ++j) {
        const size_t g = B.global0 + j * B.gstride;
        int st = hs[j].status;
        if (failed[j] && st == FB200_MODEL_OK) st = FB200_MODEL_FAILED;
        if (sh.io->status) sh.io->status[g] = st;
        if (sh.io->rows_valid) sh.io->rows_valid[g] = hs[j].rows_valid;
        if (sh.io->picard_iters) sh.io->picard_iters[g] = hs[j].picard_iters;
        if (sh.io->t_end) sh.io->t_end[g] = hs[j].t;
        steps += hs[j].rows_valid > 0 ? hs[j].rows_valid - 1 : 0;
    }
    {
        std::lock_guard<std::mutex> g(sh.err_mutex);
        sh.io->model_steps += steps;
    }
    if (dbg) fprintf(stderr, "[fb200] device %d: status scattered at %8.2f ms\n", B.device, sh.ms());
    FB_TRYB(cudaStreamSynchronize(e->copy_out));
    if (dbg) fprintf(stderr, "[fb200] device %d: rows on the host at %8.2f ms\n", B.device, sh.ms());
    return FB200_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ entry points
int plan_ensemble(const fb200_args_t* args, size_t n_models, const fb200_config_t* cfg_in, int* n_rows, int* n_cols, int* Nx, int* is_ns,
                  size_t* bad_model) {
    if (bad_model) *bad_model = 0;
    if (!args || n_models == 0) return fb_set_err(FB200_ERR_INVALID_ARGUMENT, "null argument or empty ensemble");
    fb200_config_t cfg;
    if (cfg_in) cfg = *cfg_in; else fb200_config_default(&cfg);
    Plan P;
    const int rc = make_plan(args, n_models, cfg, P, bad_model);
    if (rc != FB200_OK) return rc;
    if (n_rows) *n_rows = P.n_rows;
    if (n_cols) *n_cols = P.n_cols;
    if (Nx) *Nx = P.Nx;
    if (is_ns) *is_ns = P.is_ns ? 1 : 0;
    return FB200_OK;
}

int ensemble_launch(fb200_ensemble* e, long long n_steps) {
    FB_CUDA(cudaMemsetAsync(e->dev.queue, 0, sizeof(int), e->stream));
    FB_CUDA(cudaMemsetAsync(e->dev.progress, 0, sizeof(int) * e->n, e->stream));
    e->dev.n_steps = (n_steps < 0 || n_steps > 0x7fffffff) ? -1 : int(n_steps);
    {
        const long long most = (e->dev.n_steps < 0 || e->dev.n_steps > e->max_Nt) ? e->max_Nt : e->dev.n_steps;
        // Rounds of chunk_steps steps, then — for the last quarter of the steps, at most two chunks — rounds of half that:
        // models differ in cost, so the last wave of work items keeps only part of the CTAs busy; with shorter items at the end
        // that wave is shorter too (measured on the 10^3-model sweep: 40.40 ms with all rounds alike, 40.88 / 40.06 / 40.15 /
        // 40.04 ms with tail rounds of 3 / 8 / 10 / 12 steps; FB200_TAIL_STEPS=0: all rounds alike).
        static const int tail_env = [] { const char* c = getenv("FB200_TAIL_STEPS"); return c ? atoi(c) : -1; }();
        const long long chunk = e->chunk_steps;
        long long tail_steps = tail_env >= 0 ? tail_env : std::max<long long>(1, chunk / 2);
        long long tail_total = std::min<long long>(most / 4, 2 * chunk);
        if (tail_steps <= 0 || tail_steps >= chunk || most <= chunk) tail_total = 0, tail_steps = chunk;
        const long long n_head = (most - tail_total + chunk - 1) / chunk;           // rounds of `chunk` steps
        const long long left = std::max<long long>(0, most - n_head * chunk);        // what the short rounds have to cover
        const long long n_tail = (left + tail_steps - 1) / tail_steps;
        e->dev.chunk_steps = e->chunk_steps;
        e->dev.n_head = int(n_head);
        e->dev.tail_steps = int(tail_steps);
        e->dev.n_rounds = int(std::max<long long>(1, n_head + n_tail));
    }
    FB_CUDA(cudaEventRecord(e->ev0, e->stream));
    const int n_ctas = int(std::min<size_t>(e->n, size_t(e->dev.max_ctas)));
    cudaError_t ce = e->kb->launch_evolve(&e->dev, n_ctas, e->stream);
    if (ce != cudaSuccess) return fb_cuda_fail(ce, "fb200_evolve_kernel launch");
    FB_CUDA(cudaEventRecord(e->ev1, e->stream));
    e->last_launches = 1;
    e->timed = true;
    return FB200_OK;
}

void ensemble_free(fb200_ensemble* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    if (e->copy_in) cudaStreamSynchronize(e->copy_in);
    if (e->copy_out) cudaStreamSynchronize(e->copy_out);
    if (e->stream) {
        arena_release(e->arena, e->stream);
        arena_release(e->scratch, e->stream);
    }
    pinned_release(e->pinned_block);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    if (e->stream) cudaStreamDestroy(e->stream);
    if (e->copy_in) cudaStreamDestroy(e->copy_in);
    if (e->copy_out) cudaStreamDestroy(e->copy_out);
    delete e;
}

int build_ensembles(const fb200_args_t* args, size_t n_models, const fb200_config_t* cfg_in, const int* devices, int n_devices,
                    bool keep_initial, SweepIO* io, fb200_ensemble** out, size_t* bad_model) {
    if (bad_model) *bad_model = 0;
    if (!args || n_models == 0 || n_devices < 1 || !devices) return fb_set_err(FB200_ERR_INVALID_ARGUMENT, "null argument or empty ensemble");
    if (n_models > size_t(0x7fffffff) / 64) return fb_set_err(FB200_ERR_UNSUPPORTED, "too many models for one call");
    const bool dbg = getenv("FB200_DEBUG") != nullptr;  // FB200_DEBUG: where the time of a construction goes, on stderr
    struct DeviceGuard {  // the caller's current device is left as it was
        int dev = -1;
        DeviceGuard() { if (cudaGetDevice(&dev) != cudaSuccess) { dev = -1; cudaGetLastError(); } }
        ~DeviceGuard() { if (dev >= 0) cudaSetDevice(dev); }
    } device_guard;
    Shared sh;
    sh.t0 = std::chrono::steady_clock::now();
    if (cfg_in) sh.cfg = *cfg_in; else fb200_config_default(&sh.cfg);
    fb200_config_t& cfg = sh.cfg;
    if (io) cfg.record_F = 0;  // a one-shot sweep returns rows only
    if (cfg.n_lambdas < 0 || cfg.n_lambdas > FB200_MAX_LAMBDAS) return fb_set_err(FB200_ERR_INVALID_ARGUMENT, "n_lambdas out of range");
    if (cfg.n_passbands < 0 || cfg.n_passbands > FB200_MAX_PASSBANDS) return fb_set_err(FB200_ERR_INVALID_ARGUMENT, "n_passbands out of range");
    sh.passbands.resize(cfg.n_passbands);
    for (int p = 0; p < cfg.n_passbands; ++p) {
        std::string msg;
        const int rc = prepare_passband(cfg.passband_lambdas[p], cfg.passband_transmissions[p], cfg.passband_len[p], sh.passbands[p], msg);
        if (rc != FB200_OK) return fb_set_err(rc, "passband " + std::to_string(p) + ": " + msg);
    }
    {
        const int rc = make_plan(args, n_models, cfg, sh.P, bad_model);
        if (rc != FB200_OK) return rc;
    }
    // the band-integral table of the first model's band (a sweep shares one band; models with another ratio walk)
    if (cfg.compute_lx && !cfg.exact_lx_walk && args[0].emin > 0. && args[0].emax > args[0].emin) {
        sh.lx_r = args[0].emax / args[0].emin;
        sh.lx_table = lx_table_get(sh.lx_r, &sh.lx_doubles);
    }
    sh.keep_initial = keep_initial;
    sh.io = io;
    sh.n_total = n_models;
    if (size_t(n_devices) > n_models) n_devices = int(n_models);
    sh.n_dev = n_devices;

    // ---- companion star (--starflux): one triangle table per distinct (semi-axis, mass ratio, fill factor, starlod) ----
    if (cfg.star_flux) {
        struct Key { double semiaxis, q, fill; int lod; long long off; int ntri; };
        std::vector<Key> keys;
        sh.star_off.resize(n_models);
        sh.star_ntri.resize(n_models);
        for (size_t i = 0; i < n_models; ++i) {
            const double q = args[i].Mopt / args[i].Mx, sa = star_semiaxis_of(args[i]);
            const Key* hit = nullptr;
            for (const Key& k : keys)
                if (k.semiaxis == sa && k.q == q && k.fill == args[i].rochelobefill && k.lod == args[i].starlod) { hit = &k; break; }
            if (!hit) {
                StarGeometry g;
                std::string msg;
                const int rc = build_star_geometry(sa, q, args[i].rochelobefill, args[i].starlod, g, msg);
                if (rc != FB200_OK) {
                    if (bad_model) *bad_model = i;
                    return fb_set_err(rc, msg);
                }
                keys.push_back(Key{sa, q, args[i].rochelobefill, args[i].starlod, (long long)sh.star_table.size(), g.n_tri});
                const std::vector<double> packed = g.packed();
                sh.star_table.insert(sh.star_table.end(), packed.begin(), packed.end());
                hit = &keys.back();
            }
            sh.star_off[i] = hit->off;
            sh.star_ntri[i] = hit->ntri;
            sh.star_ntri_max = std::max(sh.star_ntri_max, hit->ntri);
        }
    }

    // ---- per-device shares, chunks, staging ----
    // upload chunks: 16-model multiples (EnsembleDev::ready); a small first one so that a streamed kernel gets to work early
    int group0 = 288, group = 576;
    if (const char* c = getenv("FB200_GROUP0")) group0 = std::max(16, atoi(c) / 16 * 16);
    if (const char* c = getenv("FB200_GROUP")) group = std::max(16, atoi(c) / 16 * 16);
    std::vector<std::unique_ptr<DeviceBuild>> builds;
    size_t max_chunk = 0;
    for (int d = 0; d < n_devices; ++d) {
        builds.emplace_back(new DeviceBuild());
        DeviceBuild& B = *builds.back();
        B.device = devices[d];
        B.global0 = size_t(d);
        B.gstride = size_t(n_devices);
        B.args.base = args + d;
        B.args.stride = size_t(n_devices);
        B.args.n = (n_models - size_t(d) + size_t(n_devices) - 1) / size_t(n_devices);
        B.chunk_begin = partition_groups(B.args.n, group0, group);
        const int n_chunks = int(B.chunk_begin.size()) - 1;
        B.chunk_left.reset(new std::atomic<int>[n_chunks]);
        for (int c = 0; c < n_chunks; ++c) {
            B.chunk_left[c].store(B.chunk_begin[c + 1] - B.chunk_begin[c]);
            max_chunk = std::max(max_chunk, size_t(B.chunk_begin[c + 1] - B.chunk_begin[c]));
        }
        auto* e = new fb200_ensemble();
        B.e = e;
        e->device = B.device;
        e->cfg = cfg;
        for (int p = 0; p < FB200_MAX_PASSBANDS; ++p) { e->cfg.passband_lambdas[p] = nullptr; e->cfg.passband_transmissions[p] = nullptr; }  // caller-owned
        e->cfg.device = B.device;
        e->n = B.args.n;
        e->Nx = sh.P.Nx;
        e->is_ns = sh.P.is_ns;
        e->n_rows = sh.P.n_rows;
        e->max_Nt = sh.P.max_Nt;
        e->n_cols = sh.P.n_cols;
        if (const char* c = getenv("FB200_CHUNK_STEPS")) e->chunk_steps = atoi(c) > 0 ? atoi(c) : 0x3fffffff;
        e->info.resize(e->n);
    }
    auto cleanup = [&] {
        for (auto& b : builds) {
            for (int s = 0; s < kSlots; ++s) {
                pinned_release(b->slots[s]);
                if (b->slot_ev[s]) { cudaSetDevice(b->device); cudaEventDestroy(b->slot_ev[s]); }
            }
            pinned_release(b->marks_block);
        }
    };
    auto destroy_all = [&] {
        for (auto& b : builds) { ensemble_free(b->e); b->e = nullptr; }
    };
    {
        // pinned staging needs a CUDA context: fail here, loudly, when there is no device
        int ndev = 0;
        cudaError_t ce = cudaGetDeviceCount(&ndev);
        if (ce != cudaSuccess || ndev == 0) {
            const std::string why = (ce != cudaSuccess) ? std::string("cudaGetDeviceCount: ") + cudaGetErrorString(ce) : std::string("no CUDA device");
            cudaGetLastError();
            for (auto& b : builds) delete b->e;
            return fb_set_err(FB200_ERR_NO_DEVICE, "no CUDA device: " + why + " (there is no CPU fallback)");
        }
        for (int d = 0; d < n_devices; ++d)
            if (devices[d] < 0 || devices[d] >= ndev) {
                for (auto& b : builds) delete b->e;
                return fb_set_err(FB200_ERR_INVALID_ARGUMENT, "device ordinal out of range");
            }
        cudaSetDevice(devices[0]);
    }
    const SlotLayout L = slot_layout(sh.P, max_chunk);
    for (auto& b : builds) {
        b->L = L;
        const int n_chunks = int(b->chunk_begin.size()) - 1;
        const int n_slots = std::min(kSlots, n_chunks);
        bool ok = true;
        for (int s = 0; s < n_slots; ++s) {
            b->slots[s] = static_cast<char*>(pinned_acquire(L.bytes));
            ok &= b->slots[s] != nullptr;
        }
        b->marks_block = pinned_acquire(sizeof(int) * (n_chunks + 1));
        ok &= b->marks_block != nullptr;
        if (!ok) {
            cleanup();
            for (auto& bb : builds) delete bb->e;
            return fb_set_err(FB200_ERR_CUDA, "pinned staging memory could not be allocated");
        }
        b->marks = static_cast<int*>(b->marks_block);
        for (int c = 0; c < n_chunks; ++c) b->marks[c] = b->chunk_begin[c + 1];
        b->marks[n_chunks] = -1;
    }
    if (dbg) fprintf(stderr, "[fb200] build: plan + staging            %8.2f ms\n", sh.ms());

    // ---- upload order: chunk c of every device before chunk c + 1 of any ----
    struct Segment { int d, c; size_t flat_begin, flat_end; };
    std::vector<Segment> segs;
    {
        size_t flat = 0;
        int max_chunks = 0;
        for (auto& b : builds) max_chunks = std::max(max_chunks, int(b->chunk_begin.size()) - 1);
        for (int c = 0; c < max_chunks; ++c)
            for (int d = 0; d < n_devices; ++d) {
                DeviceBuild& B = *builds[d];
                if (c >= int(B.chunk_begin.size()) - 1) continue;
                const size_t cnt = size_t(B.chunk_begin[c + 1] - B.chunk_begin[c]);
                segs.push_back(Segment{d, c, flat, flat + cnt});
                flat += cnt;
            }
    }

    // ---- device driver threads + host workers ----
    std::vector<std::thread> drivers;
    std::vector<int> driver_rc(n_devices, FB200_OK);
    for (int d = 0; d < n_devices; ++d)
        drivers.emplace_back([&, d] {
            driver_rc[d] = drive_device(*builds[d], sh);
            if (driver_rc[d] != FB200_OK) sh.fail(driver_rc[d], builds[d]->err, ~size_t(0));  // (keeps an earlier, more specific error)
        });

    std::atomic<size_t> next{0};
    std::atomic<double> resolve_done_ms{0.};
    unsigned nt = cfg.host_threads > 0 ? unsigned(cfg.host_threads) : HostPool::default_threads();
    if (nt > n_models) nt = unsigned(n_models);
    if (nt < 1) nt = 1;
    HostPool::instance().run(nt, [&](unsigned) {
        Resolved R;
        std::string err;
        size_t seg = 0;
        for (;;) {
            const size_t idx = next.fetch_add(1);
            if (idx >= n_models || sh.abort.load()) break;
            while (idx >= segs[seg].flat_end) ++seg;
            const Segment& S = segs[seg];
            DeviceBuild& B = *builds[S.d];
            const size_t k = idx - S.flat_begin, j = size_t(B.chunk_begin[S.c]) + k, global = B.global0 + j * B.gstride;
            while (!sh.abort.load() && S.c >= B.uploaded.load(std::memory_order_acquire) + kSlots) nap();
            if (sh.abort.load()) break;
            const int rc = resolve(B.args[j], R, err);
            if (rc != FB200_OK) {
                sh.fail(rc, err, global);
                break;
            }
            if (R.dev.Nx != sh.P.Nx || (R.dev.is_ns != 0) != sh.P.is_ns || R.dev.Nt > sh.P.max_Nt || (!R.wA.empty() && !sh.P.has_A) ||
                (!R.wB.empty() && !sh.P.has_B) || (!R.wC.empty() && !sh.P.has_C) || (R.dev.wind_dynamic && !sh.P.has_dyn)) {
                sh.fail(FB200_ERR_LOGIC, "internal: the resolved model does not match the allocation plan", global);
                break;
            }
            B.e->info[j] = R.info;
            pack_model(B, sh, S.c, k, R, global);
            if (idx + 1 == n_models) resolve_done_ms.store(sh.ms());
            B.chunk_left[S.c].fetch_sub(1, std::memory_order_acq_rel);
        }
    });
    if (dbg) fprintf(stderr, "[fb200] build: models resolved           %8.2f ms\n", sh.ms());
    for (auto& t : drivers) t.join();
    if (dbg) fprintf(stderr, "[fb200] build: devices done              %8.2f ms\n", sh.ms());
    cleanup();

    int rc = FB200_OK;
    std::string msg;
    if (sh.err_code != FB200_OK) {
        rc = sh.err_code;
        msg = sh.err_msg;
        if (bad_model && sh.err_model != ~size_t(0)) *bad_model = sh.err_model;
    } else {
        for (int d = 0; d < n_devices; ++d)
            if (driver_rc[d] != FB200_OK) { rc = builds[d]->rc != FB200_OK ? builds[d]->rc : driver_rc[d]; msg = builds[d]->err; break; }
    }
    if (rc != FB200_OK) {
        destroy_all();
        return fb_set_err(rc, msg);
    }
    if (io) {
        io->launches = n_devices;
        io->resolve_ms = resolve_done_ms.load();
        for (auto& b : builds) {
            io->device_ms_max = std::max(io->device_ms_max, b->device_ms);
            io->first_launch_ms = std::max(io->first_launch_ms, b->t_launch_ms);
            io->uploads_ms = std::max(io->uploads_ms, b->t_uploaded_ms);
        }
        destroy_all();
        io->total_ms = sh.ms();
        if (dbg) fprintf(stderr, "[fb200] build: ensembles destroyed      %8.2f ms\n", sh.ms());
    } else {
        for (int d = 0; d < n_devices; ++d) out[d] = builds[d]->e;
    }
    return FB200_OK;
}

}  // namespace fb200
