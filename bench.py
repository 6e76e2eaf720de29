#!/usr/bin/env python
"""Benchmark of the disc-evolution hot path (BASELINE.json: ensemble disc-model steps/s at Nx=1000; light curves/s).

One "step" of this bench = one pass of the hot path over one batch = evolving the whole ensemble of the workload through
all its time steps (a model-step = FreddiEvolution::step + one light-curve row).  At N=1 the workload is BASELINE
configs[1]: 10^3 BH models, alpha(40) x Cirr(25), self-irradiation, Thot=1e4, Nx=1000, 200 steps each; at N>1 every rank
gets its own 10^3 models of an N-times finer alpha grid (weak scaling, interleaved sharding, no collective on the data path).

On top of that line, at N>1 the run also does the BASELINE configuration that is quoted on that many GPUs, at full size,
and reports it as an object of the same JSON line:
    N = 8      sweep_1e5       configs[4]: 10^5 models, Mx(10) x alpha(20) x Mdisk0(25) x Cirr(20)  (the north-star target)
    N = 2, 4   sweep_1e4_wind  configs[3]: 10^4 models, Woods1996 wind + irradiation, alpha(100) x Cirr(100), 3 bands per row
    N = 1      sweep_1e4_ns    configs[2]: 10^4 neutron-star models, Bx(100) x alpha(100)
each with device-timed and construction-inclusive end-to-end throughput, the reference's CPU code on a stride sample of
the same sweep on all host cores (rank 0), the speed-up, and PARITY of the sampled models' rows against that reference.

    python bench.py --gpus N --steps K --warmup W            # the sm_100a kernels through the C ABI
    python bench.py --impl reference --gpus N ...            # the reference's own CPU code (oracle/_ref), all host cores

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'ensemble disc-model steps/sec at Nx=1000'
UNIT = 'model-steps/s'
N_ALPHA, N_CIRR = 40, 25
LAMBDAS = [3e-5, 5e-5, 8e-5]  # 3000/5000/8000 A band fluxes in every row
RTOL = 1e-10                  # BASELINE.json north_star

# Algorithmic FP64 work, SURVEY.md 8(d): fixed op counts per radial point; weights = FP64 flop (DFMA = 2, DMUL = DADD = 1)
# one call of CUDA 12.9's sm_100a math function EXECUTES on ordinary operands, measured with ncu
# (profiles/sass_flop_weights.txt, scripts/sass_flop_weights.py).
W_POW, W_EXP, W_LOG, W_DIV, W_SQRT, W_ADDMUL = 121., 29., 44., 15., 13., 1.


def algorithmic_flops(c, n_lambda):
    """SURVEY.md 8(d) evaluated on what the run really had (Ensemble.work_counters summed over models):
    ring_iters = sum over solves of (last - first) x Picard iterations, ring_steps = sum over solves of (last - first),
    row_rings = sum over rows of (last - first + 1), lx_trials = try_step calls of the Lx quadrature."""
    solve = c['ring_iters'] * (2 * W_POW + 6 * W_DIV + 10 * W_ADDMUL) + c['ring_steps'] * ((4 * W_DIV + 25 * W_ADDMUL) + (2 * W_POW + 3 * W_DIV))
    row = c['row_rings'] * (12 * W_POW + 3 * W_LOG + 2 * W_SQRT + 6 * W_DIV + 60 * W_ADDMUL) + \
        c['lx_trials'] * (6 * W_EXP + 6 * W_DIV + 60 * W_ADDMUL) + c['row_rings'] * n_lambda * (W_EXP + 2 * W_DIV + 8 * W_ADDMUL)
    return solve + row


# ---- workloads, as packed argument arrays ----
def config2_grid(n_gpus):
    return np.linspace(0.1, 1.0, N_ALPHA * n_gpus), np.geomspace(1e-5, 5e-3, N_CIRR)


SWEEP5_GRID = (np.linspace(3, 15, 10), np.linspace(0.1, 1.0, 20), np.geomspace(1e25, 1e27, 25), np.geomspace(1e-5, 5e-3, 20))
SWEEP4_GRID = (np.linspace(0.1, 1.0, 100), np.geomspace(1e-5, 5e-3, 100))
SWEEP3_GRID = (np.geomspace(10 ** 7.5, 10 ** 8.5, 100), np.linspace(0.1, 1.0, 100))


def workload(name, n_gpus, indices=None):
    from freddi_b200 import workloads as w
    if name == 'config2':
        return w.config2_sweep(*config2_grid(n_gpus), indices=indices)
    if name == 'sweep_1e5':
        return w.config5_sweep(*SWEEP5_GRID, indices=indices)
    if name == 'sweep_1e4_wind':
        return w.config4_sweep(*SWEEP4_GRID, indices=indices)
    if name == 'sweep_1e4_ns':
        return w.config3_sweep(*SWEEP3_GRID, indices=indices)
    raise KeyError(name)


def workload_size(name, n_gpus):
    return {'config2': N_ALPHA * n_gpus * N_CIRR, 'sweep_1e5': 100000, 'sweep_1e4_wind': 10000, 'sweep_1e4_ns': 10000}[name]


WORKLOAD_TEXT = {
    'config2': 'BASELINE configs[1]: BH alpha x Cirr sweep, quasistat, Thot=1e4, boundcond=Tirr, Nx=1000, 200 steps, 3 band fluxes per row',
    'sweep_1e5': 'BASELINE configs[4]: Mx(10) x alpha(20) x Mdisk0(25) x Cirr(20) = 100000 models, Nx=1000, 200 steps, 3 band fluxes per row',
    'sweep_1e4_wind': 'BASELINE configs[3]: Woods1996 wind + irradiation (Readme.md:997-1003), alpha(100) x Cirr(100) = 10000 models, Nx=1000, '
                      '200 steps, 3 band fluxes every row',
    'sweep_1e4_ns': 'BASELINE configs[2]: neutron star, freddi-ns CI command (.github/workflows/main.yml:27), Bx(100) x alpha(100) = 10000 '
                    'models, Nx=1000, 200 steps, 3 band fluxes per row, 25 + 3 columns',
}


def shard(n_models, rank, world):
    return np.arange(rank, n_models, world)  # model k -> GPU k mod G (freddi_b200.sharding.shard_indices)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, device):
        self.device = device
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.device), '--query-gpu=' + self.FIELDS,
                                          '--format=csv,noheader,nounits', '-lms', '50'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for line in self.lines:
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(smax) if smax else None,
                    reasons=sorted(reasons), samples=len(sm))


def dist_setup():
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    dist = None
    if world > 1:
        import torch.distributed as dist
    return rank, world, local, dist


# ---- the reference's CPU implementation (test infrastructure: oracle/_ref) ----
def cpu_reference(arr, cores=None, want_rows=False, n_rows=201):
    """oracle/_ref (the reference's unmodified sources) on the packed argument array `arr`: one model per thread across
    `cores` host threads.  Returns model-steps, construction and evolve seconds, and (want_rows) rows/rows_valid/picard."""
    from oracle import pyoracle
    ref = pyoracle.Ref()
    cores = cores or os.cpu_count() or 1
    out = ref.run_ensemble(arr, LAMBDAS, n_threads=cores, want_rows=want_rows, rows_per_model=n_rows if want_rows else 0)
    out['cores'] = cores
    return out


def cpu_baseline_object(out, n_sample, n_total, what):
    total_s = out['construct_s'] + out['evolve_s']
    return dict(value=out['steps'] / total_s, unit=UNIT, cores=out['cores'], kind='reference',
                sample='{} of {} models (stride-sampled over the {}), all steps of each; evolve {:.2f} s + model construction {:.2f} s, '
                       'both included'.format(n_sample, n_total, what, out['evolve_s'], out['construct_s']),
                steps=out['steps'], evolve_s=out['evolve_s'], construct_s=out['construct_s'], evolve_only_value=out['steps'] / out['evolve_s'])


def rel_err(v, g):
    """Parity rule of SURVEY.md 8(d): relative error; an exact zero, NaN or inf of the reference must be matched exactly."""
    v = np.asarray(v, dtype=float)
    g = np.asarray(g, dtype=float)
    with np.errstate(all='ignore'):
        rel = np.abs(v - g) / np.abs(g)
    rel = np.where(g == 0, np.where(v == 0, 0., np.inf), rel)
    rel = np.where(np.isnan(g) | np.isnan(v), np.where(np.isnan(g) & np.isnan(v), 0., np.inf), rel)
    rel = np.where(np.isinf(g) | np.isinf(v), np.where(v == g, 0., np.inf), rel)
    return rel


def parity_object(columns, gpu_rows, gpu_valid, gpu_picard, ref_out):
    """GPU rows of the sampled models against the reference's rows of the same models: every column of every valid row."""
    ref_rows, ref_valid, ref_picard = ref_out['rows'], ref_out['rows_valid'], ref_out['picard']
    worst = np.zeros(len(columns))
    rows_checked = 0
    for k in range(len(ref_valid)):
        n = int(min(ref_valid[k], gpu_valid[k]))
        if n > 0:
            worst = np.maximum(worst, rel_err(gpu_rows[k, :n], ref_rows[k, :n]).max(axis=0))
        rows_checked += n
    rows_valid_equal = bool((np.asarray(gpu_valid) == np.asarray(ref_valid)).all())
    picard_equal = bool((np.asarray(gpu_picard) == np.asarray(ref_picard)).all())
    t_exact = bool(all((gpu_rows[k, :int(min(ref_valid[k], gpu_valid[k])), 0] == ref_rows[k, :int(min(ref_valid[k], gpu_valid[k])), 0]).all()
                       for k in range(len(ref_valid))))
    max_rel = float(worst.max())
    return dict(models=int(len(ref_valid)), rows=int(rows_checked), rtol=RTOL, max_rel=max_rel,
                max_rel_by_column={c: float(w) for c, w in zip(columns, worst)},
                rows_valid_equal=rows_valid_equal, picard_equal=picard_equal, t_bit_exact=t_exact,
                ok=bool(max_rel <= RTOL and rows_valid_equal and picard_equal and t_exact),
                against='oracle/_ref (the reference\'s unmodified sources) on the same argument structs')


def run_reference(opts):
    rank, world, local, dist = dist_setup()
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_total = workload_size('config2', 1)
    n = min(n_total, opts.ref_models or max(8, 8 * cores))
    idx = np.linspace(0, n_total - 1, n).round().astype(int)
    arr = workload('config2', 1, idx)
    times, steps, models = [], 0, 0
    last = None
    for k in range(opts.warmup + opts.steps):
        last = cpu_reference(arr, cores)
        if k >= opts.warmup:
            times.append(last['construct_s'] + last['evolve_s'])
            steps += last['steps']
            models += n
    value = steps / sum(times)
    cb = cpu_baseline_object(last, n, n_total, 'alpha x Cirr sweep')
    cb['value'] = value
    line = dict(impl='reference', metric=METRIC, value=value, unit=UNIT, n_gpus=opts.gpus, steps=opts.steps, warmup=opts.warmup,
                ms_per_step=1e3 * sum(times) / len(times), higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f64',
                data='synthetic',
                config=dict(workload=WORKLOAD_TEXT['config2'] + ' (alpha(40) x Cirr(25))', sample=cb['sample'],
                            timed='model construction + dump/step loop of every sampled model (cpp/include/application.hpp:25-43), '
                                  'one model per host thread'),
                cpu_baseline=cb,
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                light_curves_per_s=models / sum(times))
    print(json.dumps(line), flush=True)


class Rig:
    """Per-rank context: device, distributed handles, barrier, reductions."""

    def __init__(self):
        import torch
        self.torch = torch
        self.rank, self.world, self.local, self.dist = dist_setup()
        torch.cuda.set_device(self.local)
        if self.dist is not None:
            self.dist.init_process_group('nccl', device_id=torch.device('cuda', self.local))
        # a barrier the waiting ranks sit in on the CPU (the NCCL barrier parks a spinning kernel on their GPUs, which a
        # second process using those GPUs would be time-sliced against)
        self.cpu_group = self.dist.new_group(backend='gloo') if self.dist is not None else None
        self.n_gpus = self.world if self.world > 1 else 1
        self.flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')  # > 126 MB L2

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def cpu_barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier(group=self.cpu_group)

    def reduce(self, times, work):
        t = self.torch.tensor(list(times), dtype=self.torch.float64, device='cuda')
        w = self.torch.tensor(list(work), dtype=self.torch.float64, device='cuda')
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            self.dist.all_reduce(w, op=self.dist.ReduceOp.SUM)
        return t.tolist(), w.tolist()

    def gather(self, array):
        """(world, ...) stack of one equally-shaped float64/int64 array per rank, on every rank."""
        if self.dist is None:
            return array[None]
        t = self.torch.from_numpy(np.ascontiguousarray(array)).cuda()
        parts = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(parts, t)
        return np.stack([p.cpu().numpy() for p in parts])

    def pinned(self, shape):
        return self.torch.empty(shape, dtype=self.torch.float64).pin_memory().numpy()

    def close(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()


def host_threads_per_rank():
    """Host threads this rank may use for argument resolution: the box's cores shared between the ranks of the node."""
    cores = os.cpu_count() or 1
    return max(1, cores // max(1, int(os.environ.get('LOCAL_WORLD_SIZE', '1'))))


def end_to_end_once(arr, pinned, device):
    """The call a user makes, host buffers in and out: argument structs -> freddi_b200.sweep (fb200_sweep_run: the evolve kernel
    is launched first, the host threads resolve the models into pinned staging, uploaded groups are picked up by the running
    kernel, rows of finished groups stream back) -> rows in pinned host memory.  Returns (model-steps, status arrays, stats)."""
    from freddi_b200 import sweep
    res = sweep(arr, lambdas=LAMBDAS, devices=[device], host_threads=host_threads_per_rank(), out=pinned)
    return int((res.rows_valid - 1).sum()), (res.status, res.rows_valid, res.picard_iters), res.stats


def in_process_sweep(name, n_gpus, n_rows, n_cols, check_index, check_rows):
    """The same full-size sweep from ONE process over all GPUs of the box (fb200_sweep_run with a device list: one driver thread
    and one stream set per GPU, rows scattered into one pinned buffer) — no torchrun, no torch on the path.  Run by rank 0 while
    the other ranks idle; checked bit for bit against the rows the ranks produced."""
    from freddi_b200 import sweep, pinned_empty
    n_total = workload_size(name, n_gpus)
    arr = workload(name, n_gpus, None)
    out = pinned_empty((n_total, n_rows, n_cols))
    times, stats = [], None
    for _ in range(2):  # first call: contexts on the other devices, staging, arenas
        t0 = time.perf_counter()
        res = sweep(arr, lambdas=LAMBDAS, devices=list(range(n_gpus)), out=out)
        times.append(time.perf_counter() - t0)
        stats = res.stats
    steps = int((res.rows_valid - 1).sum())
    same = bool((out[check_index].view(np.uint64) == check_rows.view(np.uint64)).all())
    return dict(value=steps / times[-1], unit=UNIT, seconds=times[-1], first_call_seconds=times[0], model_steps=steps,
                light_curves_per_s=n_total / times[-1], devices=n_gpus, stats=stats, rows_equal_the_ranks_rows_bit_for_bit=same,
                note='one Python process, fb200_sweep_run over {} devices: argument structs -> rows of all {} models in one pinned host '
                     'buffer; wall clock of the call, everything included'.format(n_gpus, n_total))


def sample_local(n_local, per_rank):
    return np.linspace(0, n_local - 1, min(per_rank, n_local)).round().astype(int)


def sweep_leg(rig, name, opts):
    """One BASELINE configuration at full size over all ranks: device-timed, end to end (construction included), CPU
    reference on a stride sample (rank 0, all host cores) with parity of the sampled rows, speed-up."""
    import ctypes as C
    from freddi_b200 import Ensemble
    from freddi_b200.args import FB200Args
    n_total = workload_size(name, rig.n_gpus)
    mine = shard(n_total, rig.rank, rig.n_gpus)
    arr = workload(name, rig.n_gpus, mine)
    n_local = len(mine)
    # warm-up: context, module load, staging slots of the size class the full sweep uses
    from freddi_b200 import sweep
    warm = sweep(workload(name, rig.n_gpus, mine[:min(n_local, 1500)]), lambdas=LAMBDAS, devices=[rig.local], host_threads=host_threads_per_rank())
    (n_rows, n_cols), columns = warm.rows.shape[1:], warm.columns
    del warm
    pinned = rig.pinned((n_local, n_rows, n_cols))
    # device-timed: inputs resident in HBM, CUDA events around the launch
    with Ensemble(arr, lambdas=LAMBDAS, device=rig.local) as e:
        e.evolve(0)
        rig.flush_buf.fill_(1)
        rig.barrier()
        e.evolve(-1)
        dev_s = e.last_evolve_ms() / 1e3
        st, valid, picard = e.status()
        launches = e.last_launches()
        steps = int((valid - 1).sum())
        collapsed = int((st == 1).sum())
    # end to end, construction included: wall clock between barriers
    rig.barrier()
    t0 = time.perf_counter()
    steps_e2e, (st2, valid2, picard2), e2e_stats = end_to_end_once(arr, pinned, rig.local)
    e2e_s = time.perf_counter() - t0
    rig.barrier()
    assert steps_e2e == steps
    # sample for parity: the same number of models from every rank's shard
    per_rank = max(1, (opts.sweep_ref_models or 256) // rig.n_gpus)
    loc = sample_local(n_local, per_rank)
    g_rows = rig.gather(pinned[loc])
    g_valid = rig.gather(valid2[loc].astype(np.int64))
    g_picard = rig.gather(picard2[loc].astype(np.int64))
    g_index = rig.gather(mine[loc].astype(np.int64))
    (t_dev, t_e2e), (steps_all, collapsed_all, models_all, launches_all) = rig.reduce(
        [dev_s, e2e_s], [steps, collapsed, n_local, launches])
    inproc = None
    if rig.n_gpus > 1 and not opts.no_in_process:
        rig.cpu_barrier()  # the other ranks wait on the CPU with idle GPUs
        if rig.rank == 0:
            try:
                inproc = in_process_sweep(name, rig.n_gpus, n_rows, n_cols, g_index.ravel(), g_rows.reshape((-1,) + g_rows.shape[2:]))
            except Exception as exc:
                inproc = dict(value=None, error=repr(exc))
        rig.cpu_barrier()
    if rig.rank != 0:
        return None
    obj = dict(workload=WORKLOAD_TEXT[name], n_gpus=rig.n_gpus, models=int(models_all), models_per_gpu=n_local,
               sharding='interleaved model index (k -> GPU k mod {}), no collective on the data path'.format(rig.n_gpus),
               value=steps_all / t_dev, unit=UNIT, evolve_s=t_dev, model_steps=int(steps_all), collapsed_models=int(collapsed_all),
               light_curves_per_s=models_all / t_dev, gpu_launches=int(launches_all),
               e2e=dict(value=steps_all / t_e2e, unit=UNIT, seconds=t_e2e, light_curves_per_s=models_all / t_e2e,
                        h2d_bytes_per_gpu=n_local * C.sizeof(FB200Args), d2h_bytes_per_gpu=int(pinned.nbytes),
                        stats_rank0=e2e_stats,
                        note='argument structs -> freddi_b200.sweep (fb200_sweep_run: streamed construction behind the running kernel, '
                             'rows of finished groups copied back meanwhile) -> rows in pinned host memory; one process per GPU; '
                             'wall clock between barriers, max over ranks'),
               in_process=inproc)
    if not opts.no_cpu_baseline:
        order = np.argsort(g_index.ravel())
        idx = g_index.ravel()[order]
        ref_out = cpu_reference(workload(name, rig.n_gpus, idx), want_rows=True, n_rows=n_rows)
        obj['cpu_baseline'] = cpu_baseline_object(ref_out, len(idx), n_total, 'sweep')
        obj['cpu_baseline']['extrapolated_full_sweep_s'] = steps_all / obj['cpu_baseline']['value']
        obj['parity'] = parity_object(columns, g_rows.reshape((-1,) + g_rows.shape[2:])[order], g_valid.ravel()[order],
                                      g_picard.ravel()[order], ref_out)
        obj['parity']['sample'] = '{} models: {} stride-sampled from the shard of every GPU'.format(len(idx), per_rank)
        obj['speedup'] = dict(device_timed=obj['value'] / obj['cpu_baseline']['value'],
                              end_to_end=obj['e2e']['value'] / obj['cpu_baseline']['value'],
                              note='against the reference on all {} host cores of this box, construction included on both sides'.format(
                                  obj['cpu_baseline']['cores']))
    return obj


def run_gpu(opts):
    import ctypes as C
    from freddi_b200 import Ensemble
    from freddi_b200.ensemble import measure_fp64_peak
    from freddi_b200.args import FB200Args

    rig = Rig()
    torch = rig.torch
    rank, local, n_gpus = rig.rank, rig.local, rig.n_gpus
    n_total = workload_size('config2', n_gpus)
    mine = shard(n_total, rank, n_gpus)
    arr = workload('config2', n_gpus, mine)
    n_local = len(mine)

    total_iters = opts.warmup + opts.steps
    ensembles = [Ensemble(arr, lambdas=LAMBDAS, device=local) for _ in range(total_iters)]
    for e in ensembles:
        e.evolve(0)  # row 0 + first-touch of every buffer: timed steps start with inputs resident in HBM
    columns = ensembles[0].columns
    n_rows, n_cols = ensembles[0].n_rows, ensembles[0].n_cols
    sampler = ClockSampler(local)
    dev_ms = []
    model_steps = models_done = 0
    counters = dict.fromkeys(Ensemble.WORK_COUNTERS, 0)
    for k, e in enumerate(ensembles):
        if k == opts.warmup:
            rig.barrier()
            sampler.start()
            t_wall0 = time.perf_counter()
        rig.flush_buf.fill_(k & 0xff)  # evict L2 between iterations
        torch.cuda.synchronize()
        e.evolve_async(-1)
        e.sync()
        if k >= opts.warmup:
            dev_ms.append(e.last_evolve_ms())
            st, valid, _ = e.status()
            model_steps += int((valid - 1).sum())
            models_done += e.n_models
            for name, v in e.work_counters().items():
                counters[name] += int(v.sum())
    rig.barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    launches = sum(e.last_launches() for e in ensembles[opts.warmup:])
    for e in ensembles:
        e.close()

    # ---- end to end through the public API with host buffers, model construction INCLUDED ----
    # One step = argument structs (host) -> freddi_b200.sweep() [fb200_sweep_run: argument resolution on the host cores, allocation,
    # H2D, evolve, rows device -> pinned host, all pipelined behind the evolve kernel].  The variant with the ensemble already
    # constructed (reset() restores the initial state from its copy in HBM) is reported beside it as e2e.resident.
    pinned = rig.pinned((n_local, n_rows, n_cols))
    e2e_times = []
    status_last = None
    for k in range(opts.warmup + opts.steps):
        rig.barrier()
        t0 = time.perf_counter()
        _steps, status_last, e2e_stats = end_to_end_once(arr, pinned, local)
        if k >= opts.warmup:
            e2e_times.append(time.perf_counter() - t0)
    h2d = n_local * C.sizeof(FB200Args)
    rows_bytes = int(pinned.nbytes)
    resident_times = []
    with Ensemble(arr, lambdas=LAMBDAS, device=local) as e:
        for k in range(opts.warmup + opts.steps):
            rig.barrier()
            t0 = time.perf_counter()
            e.reset()
            e.evolve(-1)
            e.rows(out=pinned)
            if k >= opts.warmup:
                resident_times.append(time.perf_counter() - t0)

    # ---- reduce over ranks: time = max, work = sum ----
    names = list(Ensemble.WORK_COUNTERS)
    (t_dev, t_wall, t_e2e, t_res), work = rig.reduce([sum(dev_ms) / 1e3, t_wall, sum(e2e_times), sum(resident_times)],
                                                     [model_steps, launches, models_done] + [counters[n] for n in names])
    model_steps_all, launches_all, models_all = work[:3]
    counters_all = dict(zip(names, work[3:]))

    line = None
    if rank == 0:
        value = model_steps_all / t_dev
        P = counters_all['picard_iters'] / model_steps_all
        rows_all = model_steps_all + models_all            # every model also writes row 0
        S = counters_all['lx_trials'] / max(1., counters_all['row_rings'])
        n_solve = counters_all['ring_steps'] / model_steps_all
        fl_total = algorithmic_flops(counters_all, len(LAMBDAS))
        fl = fl_total / model_steps_all
        try:
            peak = measure_fp64_peak(local)
        except Exception:
            peak = None
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        ncu = {}
        try:  # numbers of ONE launch of this workload from the committed `ncu --set full` capture (scripts/ncu_summary.py)
            ncu = json.load(open(os.path.join(ROOT, 'profiles', 'roofline_ncu.json')))
        except Exception:
            pass
        per_gpu_rate = value / n_gpus
        achieved_tf = per_gpu_rate * fl / 1e12
        launch_s = t_dev / opts.steps
        hbm_bytes_per_step = 8 * (15 + len(LAMBDAS))
        executed = None
        if ncu.get('fp64_flop_per_launch'):
            # what the hardware executed (2 DFMA + DMUL + DADD thread-instructions of one launch of THIS workload, counted by ncu)
            # over the launch duration measured live here
            ex_tf = ncu['fp64_flop_per_launch'] * (model_steps_all / opts.steps / n_gpus) / ncu['model_steps_per_launch'] / launch_s / 1e12
            executed = dict(achieved=ex_tf, unit='TFLOP/s', frac=(ex_tf / peak) if peak else None,
                            flop_per_model_step=ncu['fp64_flop_per_launch'] / ncu['model_steps_per_launch'], source=ncu.get('source'))
        roofline = dict(bound='fp64', achieved=achieved_tf, peak=peak, unit='TFLOP/s',
                        frac=(achieved_tf / peak) if peak else None, traffic=ncu.get('dram_bytes_per_launch'),
                        peak_source='DFMA micro-kernel measured in this run (MEASURED_PEAKS.json has no FP64 entry)',
                        flop_model='ALGORITHMIC: the reference\'s operation count (SURVEY.md 8d) on the rings and iterations this run really had '
                                   '(work counters of the kernels: {:.1f} unknowns per solve, P={:.2f} Picard iterations per step, {:.1f} hot rings per '
                                   'row, S={:.2f} quadrature trial steps per ring and row — the trial steps the kernel executes: rings it prunes are '
                                   'not counted) x flop per call measured with ncu (profiles/sass_flop_weights.txt: pow {:.0f}, exp {:.0f}, log {:.0f}, '
                                   'div {:.0f}, sqrt {:.0f}).  The kernel reaches the same results with fewer operations (incremental K update: '
                                   '{:.3f} pow per ring and iteration instead of 2; the band integral of Lx read from a table of the adaptive walk\'s own '
                                   'result, csrc/lx_table.h, instead of S trial steps per ring), so this fraction is NOT pipe utilisation: see '
                                   '`executed` and `fp64_pipe_active`.'.format(n_solve, P, counters_all['row_rings'] / rows_all, S, W_POW, W_EXP, W_LOG, W_DIV, W_SQRT,
                                                                counters_all['pow_evals'] / max(1., counters_all['ring_iters'])),
                        flops_per_model_step=fl, executed=executed, fp64_pipe_active=ncu.get('fp64_pipe_active_pct'),
                        issue_active=ncu.get('issue_active_pct'),
                        hbm=dict(achieved=per_gpu_rate * hbm_bytes_per_step / 1e9, peak=peaks.get('hbm_gbs'), unit='GB/s',
                                 note='algorithmic HBM traffic is one {}-byte row per model-step: irrelevant to this path'.format(
                                     hbm_bytes_per_step)))
        cpu = parity = None
        if n_gpus == 1 and not opts.no_cpu_baseline:
            try:
                cores = os.cpu_count() or 1
                n = min(n_total, opts.ref_models or max(8, 8 * cores))
                idx = np.linspace(0, n_total - 1, n).round().astype(int)
                ref_out = cpu_reference(workload('config2', 1, idx), cores, want_rows=True, n_rows=n_rows)
                cpu = cpu_baseline_object(ref_out, n, n_total, 'alpha x Cirr sweep')
                st, valid, picard = status_last
                parity = parity_object(columns, pinned[idx], valid[idx], picard[idx], ref_out)
                parity['sample'] = 'the {} models of the cpu_baseline sample, rows of the last end-to-end iteration'.format(n)
            except Exception as exc:  # the checker library is test infrastructure; the bench line survives without it
                cpu = dict(value=None, unit=UNIT, cores=os.cpu_count(), kind='reference', sample='unavailable: {}'.format(exc))
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=n_gpus, steps=opts.steps, warmup=opts.warmup,
                    ms_per_step=1e3 * t_dev / opts.steps, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f64',
                    data='synthetic',
                    config=dict(workload=WORKLOAD_TEXT['config2'] + ': alpha({}) x Cirr({}) = {} models/GPU'.format(
                        N_ALPHA * n_gpus, N_CIRR, n_local),
                                l2='flushed (256 MB write) before every timed iteration',
                                sharding='interleaved model index, no collective on the data path',
                                wall_ms_per_step=1e3 * t_wall / opts.steps),
                    roofline=roofline, cpu_baseline=cpu, parity=parity, clocks=clocks,
                    e2e=dict(value=model_steps_all / t_e2e, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=rows_bytes,
                             iter_ms=[round(1e3 * x, 2) for x in e2e_times],
                             note='per step: argument structs (host) -> freddi_b200.sweep() = fb200_sweep_run [kernel launched first; host '
                                  'threads resolve the models into pinned staging; H2D per group, picked up by the running kernel; rows of '
                                  'finished groups device -> pinned host meanwhile]; wall clock; model construction INCLUDED, as in the '
                                  'reference arm.  h2d_bytes_per_step counts the argument structs the call takes (the resolved h, F, '
                                  'constants it uploads: h2d_resolved_bytes_per_step)',
                             h2d_resolved_bytes_per_step=n_local * (2 * 1000 * 8 + 900),
                             stats_last=e2e_stats, host_threads=host_threads_per_rank(),
                             resident=dict(value=model_steps_all / t_res, unit=UNIT, iter_ms=[round(1e3 * x, 2) for x in resident_times],
                                           h2d_bytes_per_step=0,
                                           note='ensemble constructed once; per step reset() [initial F + state from their copy in HBM] + '
                                                'evolve() + rows() [device -> pinned host]')),
                    gpu_launches=int(launches_all), light_curves_per_s=models_all / t_dev,
                    model_steps_per_iteration=model_steps_all / opts.steps)

    # ---- the BASELINE configuration quoted on this many GPUs, at full size ----
    sweep_name = {1: 'sweep_1e4_ns', 8: 'sweep_1e5', 2: 'sweep_1e4_wind', 4: 'sweep_1e4_wind'}.get(n_gpus)
    if opts.sweep:
        sweep_name = opts.sweep
    if sweep_name and sweep_name != 'none':
        obj = sweep_leg(rig, sweep_name, opts)
        if rank == 0:
            line[sweep_name] = obj
    if rank == 0:
        print(json.dumps(line), flush=True)
    rig.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--ref-models', type=int, default=None, help='models in the CPU sample (default 8 x cores, max 1000)')
    ap.add_argument('--sweep-ref-models', type=int, default=None, help='models in the CPU sample of the full-size sweep (default 256)')
    ap.add_argument('--sweep', default=None, choices=['sweep_1e5', 'sweep_1e4_wind', 'sweep_1e4_ns', 'none'],
                    help='full-size configuration to add to the line (default: sweep_1e5 at 8 GPUs, sweep_1e4_wind at 2/4, sweep_1e4_ns at 1)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-in-process', action='store_true', help='skip the one-process multi-GPU run of the full-size sweep (N > 1)')
    opts = ap.parse_args()
    if opts.impl == 'reference':
        run_reference(opts)
    else:
        run_gpu(opts)


if __name__ == '__main__':
    main()
