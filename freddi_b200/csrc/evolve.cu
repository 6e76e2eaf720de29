// The persistent ensemble kernel: each CTA pulls model indices from a device-side queue, stages the
// model's FP64 state (h, F and the grid-derived constants) into shared memory once, and runs the whole
// multi-step loop there: FreddiEvolution::step (cpp/src/freddi_evolution.cpp:17-29) + one light-curve row
// (cpp/src/output.cpp:197-233) per step.  Only rows (and optional F snapshots) go back to HBM.
#include "gpu_exec.cuh"

namespace fb200 {

template <bool kExtras>
__global__ void __launch_bounds__(kNT, kCtasPerSm) fb200_evolve_kernel(const EnsembleDev E) {
    GpuArrays S = carve_arrays(E.cta_scratch + size_t(blockIdx.x) * SmemLayout::kScratchDoublesPerCta, E.h);
    __shared__ int s_model;

    GpuEx ex;
    ex.init();
    ex.set_nx(E.Nx);
    const OutCfg& cfg = stage_cfg(E);
#ifdef FB_PROFILE
    for (int p = 0; p < 16; ++p) ex.prof[p] = 0;
    ex.prof_last = clock64();
#endif
    const int tid = threadIdx.x;
    const int Nx = E.Nx;
    const int ncol = cfg.n_cols, nrows = cfg.n_rows;

    // Work items are (round, model) pairs in round-major order: a model runs chunk_steps steps per item and goes
    // back to the queue, so that the tail of the launch is a fraction of a model's run instead of a whole one.
    // (A streamed sweep keeps this order: the first round simply follows the uploads, which stay ahead of it — a round of
    // 10^3 models is 6 ms of work for the GPU, their resolution and upload 1 ms for the host.  Numbering the items upload
    // group by upload group instead made a group of fewer models than CTAs run its rounds in lock step: measured 1 ms lost
    // per 10^3-model sweep.)
    // Item (r, m) needs item (r-1, m) finished: it was handed out earlier (the queue is ordered), to a CTA that is
    // resident (grid <= resident capacity) and never waits on a later item, so the wait below cannot deadlock.
    const int n_items = E.n_rounds * E.n_models;
    __shared__ int s_abort, s_round;
    for (;;) {
        __syncthreads();  // previous item fully done with shared memory and s_model
        if (tid == 0) {
            const int item = atomicAdd(E.queue, 1);
            int model = -1, round = 0;
            if (item < n_items) {
                round = item / E.n_models;
                model = item - round * E.n_models;
                s_abort = 0;
#pragma unroll
                for (int c = 0; c < FB_CNT_N; ++c) GpuEx::counters()[c] = 0;  // bookkeeping of this work item (GpuEx::count)
                const unsigned long long t0 = fb_globaltimer();
                if (E.ready && *reinterpret_cast<const volatile int*>(E.ready) < 0) s_abort = 2;  // the host withdrew the launch
                if (round == 0 && E.ready && !s_abort) {
                    // streamed construction: wait until the model has been uploaded (the host bumps `ready`, the number of leading
                    // models in HBM, on the copy stream behind each chunk's arrays); -1 = the host gave up (an argument error in a
                    // later model)
                    const volatile int* rd = E.ready;
                    for (;;) {
                        const int have = *rd;
                        if (have < 0) { s_abort = 2; break; }
                        if (have > model) break;
                        __nanosleep(200);
                        if (fb_globaltimer() - t0 > (unsigned long long)E.wait_timeout_ns) {
                            atomicExch(E.failed + model, 1);
                            s_abort = 1;
                            break;
                        }
                    }
                    __threadfence();
                }
                if (round > 0 && !s_abort) {
                    // Wait for the producer of the previous round, by elapsed time (%globaltimer, ns).  A producer that never
                    // publishes (it can only be a CTA that died) must not hang the launch: after wait_timeout_ns the model is
                    // flagged in E.failed — a flag of its own, set with atomicExch: the state record belongs to the producer, which
                    // may still be writing it — and every later item of that model gives up at once.
                    const volatile int* p = E.progress + model;
                    const volatile int* f = E.failed + model;
                    while (*p < round) {
                        if (*f) { s_abort = 1; break; }
                        if (E.ready && *reinterpret_cast<const volatile int*>(E.ready) < 0) { s_abort = 2; break; }  // withdrawn: its producer may have left
                        __nanosleep(100);
                        if (fb_globaltimer() - t0 > (unsigned long long)E.wait_timeout_ns) {
                            atomicExch(E.failed + model, 1);
                            s_abort = 1;
                            break;
                        }
                    }
                    __threadfence();
                }
            }
            s_model = model;
            s_round = round;
        }
        __syncthreads();
        const int model = s_model;
        if (model < 0) break;
        const int round = s_round;
        if (s_abort == 2) break;  // the host withdrew the launch
        if (s_abort) continue;    // nothing of this model is touched: no state, no progress

        ModelState st;
        {
            const long long* src = reinterpret_cast<const long long*>(E.state + model);
            long long* dst = reinterpret_cast<long long*>(&st);
#pragma unroll
            for (int w = 0; w < int(sizeof(ModelState) / 8); ++w) dst[w] = __ldcg(src + w);
        }
        if (round == 0 && __ldcg(E.failed + model)) st.status = FB200_MODEL_FAILED;  // flagged in an earlier launch: stays failed
        bool active = (st.status == FB200_MODEL_OK);
        if (active && round == 0) {
            const int Nt_m = E.models[model].Nt;
            int target = Nt_m;
            if (E.n_steps >= 0 && (long long)st.i_t + E.n_steps < (long long)Nt_m) target = st.i_t + E.n_steps;
            st.target_i_t = target;
        }
        if (active && st.row0_done && st.i_t >= st.target_i_t) active = false;  // nothing left for this model in this launch
        if (active) {
            if (tid == 0) GpuEx::wind_update() = WindRecord{st.wind_Mdot, st.wind_first, st.wind_last};
            S.h = E.h + size_t(model) * Nx;
            load_model(E, model, S, ex, E.F + size_t(model) * Nx);
            const fb200_model_dev_t& M = *smem_model();

            DiscEngine<GpuEx, kExtras> eng(ex, M, cfg, S, st);
            eng.init_points();
            if (st.k_valid) {  // the closure of the stored state: results do not depend on where a run is cut into work items
                for (int i = tid; i < Nx; i += kNT) S.s_K[i] = __ldcg(E.K + size_t(model) * Nx + i);
                eng.k_valid = true;
                __syncthreads();
            }

            double* rows = E.rows + size_t(model) * nrows * ncol;
            double* Fsnap = E.Fsnap ? E.Fsnap + size_t(model) * nrows * Nx : nullptr;
            int* bounds = E.bounds ? E.bounds + size_t(model) * nrows * 2 : nullptr;

            WindRecord* wrec = (E.wind_rec && M.wind_dynamic) ? E.wind_rec + size_t(model) * nrows : nullptr;
            if (!eng.st.row0_done) {
                // the state at t0 is dumped before the first step (cpp/include/application.hpp:25-29)
                eng.structure_and_row(false, rows, Fsnap, bounds);
                eng.st.row0_done = 1;
                eng.st.rows_valid = 1;
                // the wind arrays at t0 are what the wind's constructor left: its own update() on the initial state, over
                // [first, Nx - 1] — ShieldsOscil1986 has none (T12 of SURVEY.md 9): an empty range
                if (wrec && tid == 0) {
                    const bool ctor_update = (M.windtype != FB200_WIND_SHIELDSOSCIL1986);
                    wrec[0] = WindRecord{M.wind_ctor_Mdot, ctor_update ? M.first0 : 1, ctor_update ? Nx - 1 : 0};
                }
            }
            int todo = eng.st.target_i_t - eng.st.i_t;
            const int chunk = (round < E.n_head) ? E.chunk_steps : E.tail_steps;
            if (todo > chunk) todo = chunk;
            for (int s = 0; s < todo; ++s) {
                const int r = eng.st.i_t + 1;
                const bool ok = eng.step_and_row(rows + size_t(r) * ncol, Fsnap ? Fsnap + size_t(r) * Nx : nullptr,
                                                 bounds ? bounds + size_t(r) * 2 : nullptr);
                if (!ok) break;
                eng.st.rows_valid = r + 1;
                if (wrec && tid == 0) wrec[r] = GpuEx::wind_update();
            }
            // write the state back for the next work item / evolve call / the read-out kernels
            __syncthreads();
            for (int i = tid; i < Nx; i += kNT) E.F[size_t(model) * Nx + i] = S.F[i];
            if (eng.k_valid)
                for (int i = tid; i < Nx; i += kNT) E.K[size_t(model) * Nx + i] = S.s_K[i];
            st = eng.st;
            st.k_valid = eng.k_valid ? 1 : 0;
        }
        if (tid == 0 && (active || round == 0)) {
            // Only what a work item changes is written back.  The bookkeeping fields are not carried through the step loop in the
            // threads' registers: thread 0 owns them in shared memory (GpuEx::count) and adds them to the record here (the atomics
            // of the other threads precede the barrier above).
            ModelState* g = E.state + model;
            g->t = st.t; g->F_in = st.F_in; g->Mdot_in_prev = st.Mdot_in_prev;
            g->first = st.first; g->last = st.last; g->i_t = st.i_t; g->status = st.status; g->rows_valid = st.rows_valid;
            g->row0_done = st.row0_done; g->target_i_t = st.target_i_t; g->k_valid = st.k_valid;
            if (active) {
                const long long* cnt = GpuEx::counters();
                // (read through L2 like the rest of the record: another SM wrote it in the model's previous round)
                g->picard_iters = __ldcg(&g->picard_iters) + cnt[FB_CNT_PICARD];
                g->lx_trials = __ldcg(&g->lx_trials) + cnt[FB_CNT_LX_TRIALS];
                g->ring_steps = __ldcg(&g->ring_steps) + cnt[FB_CNT_RING_STEPS];
                g->ring_iters = __ldcg(&g->ring_iters) + cnt[FB_CNT_RING_ITERS];
                g->row_rings = __ldcg(&g->row_rings) + cnt[FB_CNT_ROW_RINGS];
                g->pow_evals = __ldcg(&g->pow_evals) + cnt[FB_CNT_POW_EVALS] + cnt[FB_CNT_POW_FALLBACKS];
                const WindRecord w = GpuEx::wind_update();
                g->wind_Mdot = w.Mdot; g->wind_first = w.first; g->wind_last = w.last;
            }
        }
        __threadfence();   // F, rows and state are visible device-wide before the round is published
        __syncthreads();
        if (tid == 0) {
            atomicExch(E.progress + model, round + 1);
            if (E.block_count && round == E.n_rounds - 1) {
                // last work item of the model: when it is the last one of its read-back block, tell the host (the fence above
                // ordered this CTA's rows before the count; the counts of the other CTAs were ordered the same way)
                const int blk = model / kDoneBlock;
                const int size = (E.n_models - blk * kDoneBlock < kDoneBlock) ? E.n_models - blk * kDoneBlock : kDoneBlock;
                if (atomicAdd(E.block_count + blk, 1) == size - 1) {
                    __threadfence_system();
                    *reinterpret_cast<volatile int*>(E.block_done + blk) = 1;
                }
            }
        }
        ex.tick(7);
    }
#ifdef FB_PROFILE
    if (tid == 0 && E.prof)
        for (int p = 0; p < 16; ++p) atomicAdd(reinterpret_cast<unsigned long long*>(E.prof + p), (unsigned long long)ex.prof[p]);
#endif
}

size_t evolve_smem_bytes() { return SmemLayout::bytes(); }
int evolve_resident_ctas_per_sm() {
    int n = 0, m = 0;
    cudaFuncSetAttribute(fb200_evolve_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SmemLayout::bytes()));
    cudaFuncSetAttribute(fb200_evolve_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SmemLayout::bytes()));
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fb200_evolve_kernel<false>, kNT, SmemLayout::bytes()) != cudaSuccess) return -1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&m, fb200_evolve_kernel<true>, kNT, SmemLayout::bytes()) != cudaSuccess) return -1;
    return n < m ? n : m;
}
size_t evolve_scratch_doubles_per_cta() { return SmemLayout::scratch_doubles_per_cta(); }
int evolve_ctas_per_sm() { return kCtasPerSm; }

cudaError_t launch_evolve(const EnsembleDev& E, int n_ctas, cudaStream_t stream) {
    // per device and cheap: set on every launch so that several devices in one process all get it
    const bool extras = E.cfg.star_flux || E.cfg.n_passbands > 0;  // passband / companion-star columns: the second instantiation
    cudaError_t e = extras ? cudaFuncSetAttribute(fb200_evolve_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SmemLayout::bytes()))
                           : cudaFuncSetAttribute(fb200_evolve_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SmemLayout::bytes()));
    if (e != cudaSuccess) return e;
    if (n_ctas > E.max_ctas) n_ctas = E.max_ctas;
    if (extras) fb200_evolve_kernel<true><<<n_ctas, kNT, SmemLayout::bytes(), stream>>>(E);
    else fb200_evolve_kernel<false><<<n_ctas, kNT, SmemLayout::bytes(), stream>>>(E);
    return cudaGetLastError();
}

}  // namespace fb200

// Exports with plain-C names, so that construct.cu / capi.cu can hold one table of kernel builds (ensemble_impl.hpp,
// KernelBuild): this file is compiled once per build — the main one (Nx <= 1024, 256 threads per model), the large-grid one
// (evolve_big.cu) and the small-grid ones (evolve_s128.cu: one warp per model, evolve_s256.cu: two warps per model) — each in
// its own namespace.  EnsembleDev has the same layout in all builds (nothing in it depends on FB200_NX_MAX).
#ifndef FB_VARIANT
#define FB_VARIANT main
#endif
#define FB_CAT3_(a, b, c) a##b##c
#define FB_CAT3(a, b, c) FB_CAT3_(a, b, c)
#define FB_VNAME(name) FB_CAT3(fb200_, FB_VARIANT, name)
extern "C" {
size_t FB_VNAME(_evolve_smem_bytes)() { return fb200::evolve_smem_bytes(); }
size_t FB_VNAME(_evolve_scratch_doubles_per_cta)() { return fb200::evolve_scratch_doubles_per_cta(); }
int FB_VNAME(_evolve_ctas_per_sm)() { return fb200::evolve_ctas_per_sm(); }
int FB_VNAME(_evolve_resident_ctas_per_sm)() { return fb200::evolve_resident_ctas_per_sm(); }
int FB_VNAME(_evolve_threads)() { return fb200::kNT; }
cudaError_t FB_VNAME(_launch_evolve)(const void* E, int n_ctas, cudaStream_t stream) {
    return fb200::launch_evolve(*static_cast<const fb200::EnsembleDev*>(E), n_ctas, stream);
}
}
