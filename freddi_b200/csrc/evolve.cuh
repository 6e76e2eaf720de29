// Device-side view of an ensemble and the launch interface of the evolve kernel (evolve.cu).
#pragma once
#include <cuda_runtime.h>

#include "disc_engine.h"

namespace fb200 {

constexpr int kDoneBlock = 64;  // models per read-back block of a streamed sweep

struct EnsembleDev {
    int n_models;
    int Nx;            // every model of an ensemble has the same Nx
    int n_steps;       // steps requested by this launch (clipped per model at Nt - i_t)
    OutCfg cfg;
    const fb200_model_dev_t* models;  // [n_models]
    ModelState* state;                // [n_models]
    const double* h;                  // [n_models][Nx]
    double* F;                        // [n_models][Nx] current state
    double* K;                        // [n_models][Nx] K of the last Picard iteration, carried between work items with F
    const double* wA;                 // [n_models][Nx] or null: static wind A
    const double* wB;                 // [n_models][Nx] or null
    const double* wCs;                // [n_models][Nx] or null: static wind C + d2Fmagn_dh2
    const double* Fmagn;              // [n_models][Nx] or null
    const double* dFmagn_dh;          // [n_models][Nx] or null
    const double* d2Fmagn_dh2;        // [n_models][Nx] or null (read-out only; the solver uses wCs)
    double* rows;                     // [n_models][n_rows][n_cols]
    double* Fsnap;                    // [n_models][n_rows][Nx] or null (record_F)
    int* bounds;                      // [n_models][n_rows][2] or null (first,last per row; record_F)
    int* queue;                       // work-queue counter
    int* progress;                    // [n_models] rounds completed in this launch (chunked scheduling)
    int* failed;                      // [n_models] 1: a work item of the model gave up waiting for its producer (kept across launches)
    long long wait_timeout_ns;        // how long a work item waits for the previous round of its model before giving up
    WindRecord* wind_rec;             // [n_models][n_rows] or null: arguments of the wind update() behind every row (dynamic winds)
    int chunk_steps;                  // steps per work item; a model is handed back to the queue after that many
    int n_head, tail_steps;           // rounds >= n_head are short (tail_steps per work item): the last wave of the launch, in which
                                      // most CTAs idle, then lasts half as long
    int n_rounds;                     // work items per model in this launch
    double* cta_scratch;              // [max_ctas][scratch_doubles_per_cta()] once-per-step arrays of the model a CTA holds
    int max_ctas;                     // CTAs the scratch was sized for (2 per SM)
    long long* prof;                  // [16] per-phase cycle sums (FB_PROFILE builds only), else null
    double* star_T;                   // [max_ctas][star_ntri_max] companion-star triangle temperatures (--starflux), else null
    int star_ntri_max;
    int _pad_star;
    // Streamed construction (fb200_sweep_run; null otherwise): the kernel is launched before the models are in HBM.
    // ready: number of leading models uploaded so far, written by the copy stream behind each upload chunk's arrays (-1: give
    //        up); the first work item of a model waits until it covers the model.  Chunks are multiples of 16 models, so that
    //        no 128-byte line is shared by two uploads (a line fetched into L1 with a neighbour's tail would otherwise be stale).
    // block_count[b]: models of read-back block b (kDoneBlock consecutive models) whose last work item is finished;
    //        block_done[b] (host-mapped, pinned): set to 1 by the CTA that finishes the last one — the host then reads the block's
    //        rows back while the kernel still runs (blocks finish roughly in order: in the last round of a group the work items
    //        are handed out in model order), so that only the last block's copy is left when the kernel ends.
    const int* ready;
    int* block_count;
    int* block_done;
};

size_t evolve_smem_bytes();
size_t evolve_scratch_doubles_per_cta();
int evolve_ctas_per_sm();
int evolve_resident_ctas_per_sm();  // what the occupancy calculator grants on the current device
// Launches the persistent evolve kernel on `stream`; returns cudaGetLastError().
cudaError_t launch_evolve(const EnsembleDev& E, int n_ctas, cudaStream_t stream);

// FP64 DFMA peak of the current device in TFLOP/s (fields.cu)
cudaError_t measure_fp64_peak(double* tflops);

// policy_trapz on caller data with one CTA (fields.cu): n <= FB200_NX_MAX points, scratch_dev = one CTA scratch slot
cudaError_t selftest_trapz(const double* x_dev, const double* y_dev, int n, int first, int last, double* out_dev, double* scratch_dev,
                           cudaStream_t stream);

// field / flux read-out kernels (fields.cu)
cudaError_t launch_field(const EnsembleDev& E, int field, long long row0, int n_sel, int pad_nan, double* out, cudaStream_t stream);
cudaError_t launch_flux(const EnsembleDev& E, int region, long long row0, int n_sel, const double* lambdas_dev, int n_lambdas,
                        const double* phases_dev, double* out, cudaStream_t stream);
cudaError_t launch_mdot_wind(const EnsembleDev& E, long long row, double* out, cudaStream_t stream);

}  // namespace fb200
