// Internal to the library (capi.cu, construct.cu, sharded.cu): the ensemble object behind fb200_ensemble_t and the pieces
// its construction is made of.  Nothing here is part of the C ABI.
//
// Construction is a pipeline, so that a sweep's host work hides behind the GPU:
//   plan      cheap facts of the whole ensemble from the argument structs alone (Nx, class, rows, which optional arrays
//             exist), then ONE device allocation per GPU (taken from a per-device cache that outlives the ensemble) carved
//             into the arrays;
//   fill      persistent host workers resolve the models in upload order (their per-thread memos stay warm) straight into
//             pinned staging slots, chunk by chunk; the device's driver thread uploads a chunk with cudaMemcpyAsync as soon
//             as its last model is resolved, while the workers are already in the next chunks;
//   stream    (fb200_sweep_run) the evolve kernel is launched BEFORE the fill: its work items wait for the upload watermark
//             of their scheduling group, and the rows of a finished group are read back while later groups still run.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <functional>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <vector>

#include "evolve.cuh"
#include "resolve.hpp"

namespace fb200 {
// One compilation of the kernels (evolve.cu / fields.cu): the main build (Nx <= 1024, 256 threads per model), the small-grid
// builds (s128: one warp per model, s256: two warps, s512: four) and the large-grid build (big: arrays in an L2-resident scratch slot).
struct KernelBuild {
    const char* name;
    int nx_max;
    size_t (*smem_bytes)();
    size_t (*scratch_doubles_per_cta)();
    int (*ctas_per_sm)();
    int (*resident_ctas_per_sm)();
    int (*threads)();
    cudaError_t (*launch_evolve)(const void* E, int n_ctas, cudaStream_t stream);
    cudaError_t (*launch_field)(const void* E, int field, long long row0, int n_sel, int pad_nan, double* out, cudaStream_t stream);
    cudaError_t (*launch_flux)(const void* E, int region, long long row0, int n_sel, const double* lambdas_dev, int n_lambdas,
                               const double* phases_dev, double* out, cudaStream_t stream);
    cudaError_t (*launch_mdot_wind)(const void* E, long long row, double* out, cudaStream_t stream);
};
constexpr int kKernelBuilds = 5;
const KernelBuild& kernel_build(int index);        // 0 s128, 1 s256, 2 s512, 3 main, 4 big
int kernel_build_index_for(int Nx);                // the smallest build that holds Nx (FB200_BUILD=main: no small-grid builds)
}  // namespace fb200

struct fb200_ensemble {
    int device = 0;
    const fb200::KernelBuild* kb = nullptr;
    cudaStream_t stream = nullptr;      // evolve / read-out
    cudaStream_t copy_in = nullptr;     // H2D of the construction pipeline
    cudaStream_t copy_out = nullptr;    // D2H of finished groups (streamed sweep)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    size_t n = 0;
    int Nx = 0, n_rows = 0, n_cols = 0, sm_count = 0, max_Nt = 0, chunk_steps = 25;
    bool is_ns = false;
    fb200_config_t cfg{};
    std::vector<fb200_model_info_t> info;
    fb200::EnsembleDev dev{};
    void* arena = nullptr;             // the one device allocation the arrays are carved from (per-device cache)
    size_t arena_bytes = 0;
    int last_launches = 0;
    bool timed = false;
    double* scratch = nullptr;         // read-out staging, grown on demand (own allocation)
    size_t scratch_len = 0;
    // device copies of the initial conditions (fb200_ensemble_reset is a device-to-device copy); null for a one-shot sweep
    double* F0_dev = nullptr;
    fb200::ModelState* state0_dev = nullptr;
    int* ready_dev = nullptr;          // upload watermark (device)
    int* block_done_host = nullptr;    // [ceil(n / kDoneBlock)] mapped pinned flags of the read-back blocks
    void* pinned_block = nullptr;      // what block_done_host was cut from
};

namespace fb200 {

// error plumbing shared by the translation units (thread-local message behind fb200_last_error)
int fb_set_err(int code, const std::string& msg);
int fb_cuda_fail(cudaError_t e, const char* what);
const std::string& fb_err_msg();
#define FB_CUDA(call)                                                  \
    do {                                                               \
        cudaError_t _e = (call);                                       \
        if (_e != cudaSuccess) return fb200::fb_cuda_fail(_e, #call);  \
    } while (0)

// args[i * stride]: lets a shard view every G-th model of the caller's array without copying it
struct ArgView {
    const fb200_args_t* base = nullptr;
    size_t stride = 1, n = 0;
    const fb200_args_t& operator[](size_t i) const { return base[i * stride]; }
};

// Persistent host workers shared by every construction of the process (their thread-local resolver memos stay warm between
// calls).  run(n, body) runs body(worker_index) on n workers — the caller is worker 0 — and returns when all are done.
class HostPool {
public:
    static HostPool& instance();
    void run(unsigned n_workers, const std::function<void(unsigned)>& body);
    static unsigned default_threads();

private:
    HostPool() = default;
    struct Impl;
    Impl* impl_ = nullptr;
};

// Pinned (portable, mapped) host blocks, cached for the life of the process: cudaHostAlloc costs ~0.2 ms per MB.
void* pinned_acquire(size_t bytes);   // nullptr when pinned memory cannot be had
void pinned_release(void* block);
// Device memory from the library's stream-ordered pool of the device (kept across ensembles; no device synchronisation on
// either call).  `stream`: an idle stream of the device (acquire) / the stream the last use was ordered on (release).
int arena_acquire(int device, size_t bytes, cudaStream_t stream, void** p);
void arena_release(void* p, cudaStream_t stream);
void caches_trim();                   // gives back what the pools and the pinned cache hold idle
}  // namespace fb200
namespace fb200 {
extern std::shared_mutex g_upload_window;  // construct.cu: shared by streamed sweeps during their upload window, exclusive for allocations

// Where a streamed sweep delivers its results (all host pointers; rows may be pageable, pinned is faster).
struct SweepIO {
    double* rows = nullptr;           // [n_models][n_rows][n_cols] of the caller, models in the caller's order
    int32_t* status = nullptr;        // [n_models] or null
    int32_t* rows_valid = nullptr;    // [n_models] or null
    int64_t* picard_iters = nullptr;  // [n_models] or null
    double* t_end = nullptr;          // [n_models] or null: t of the state each model stopped at [s]
    // filled in by the call
    double device_ms_max = 0.;        // longest evolve launch over the devices (includes its waits for uploads)
    double first_launch_ms = 0.;      // entry -> the last device's kernel launched
    double resolve_ms = 0.;           // entry -> the last model resolved
    double uploads_ms = 0.;           // entry -> the last upload complete
    double total_ms = 0.;
    long long model_steps = 0;        // sum over models of rows_valid - 1
    int launches = 0;
};

// Builds one ensemble per entry of `devices` (a device may be listed more than once) over args[k], k = d, d + G, d + 2G, ...
// keep_initial: keep device copies of the initial state for fb200_ensemble_reset.  io != null: streamed sweep — evolve all
// models to their end and deliver rows / status into io; the ensembles are destroyed before returning and out may be null.
int build_ensembles(const fb200_args_t* args, size_t n_models, const fb200_config_t* cfg_in, const int* devices, int n_devices,
                    bool keep_initial, SweepIO* io, fb200_ensemble** out, size_t* bad_model);
// rows per model (max Nt + 1), columns per row and the shared Nx / class of an argument array (host only, no resolution)
int plan_ensemble(const fb200_args_t* args, size_t n_models, const fb200_config_t* cfg, int* n_rows, int* n_cols, int* Nx, int* is_ns,
                  size_t* bad_model);

int ensemble_launch(fb200_ensemble* e, long long n_steps);   // enqueue one evolve launch on e->stream (between ev0 and ev1)
void ensemble_free(fb200_ensemble* e);

}  // namespace fb200
