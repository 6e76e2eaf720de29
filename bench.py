#!/usr/bin/env python
"""Benchmark of the disc-evolution hot path (BASELINE.json: ensemble disc-model steps/s at Nx=1000).

One "step" of this bench = one pass of the hot path over one batch = evolving the whole ensemble of the
workload through all its time steps (a model-step = FreddiEvolution::step + one light-curve row).  At N=1 the
workload is BASELINE configs[1]: 10^3 BH models, alpha(40) x Cirr(25), self-irradiation, Thot=1e4, Nx=1000,
200 steps each; at N>1 every rank gets its own 10^3 models of an N-times finer alpha grid (weak scaling,
interleaved sharding, no collective on the data path).

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

# Algorithmic FP64 work, SURVEY.md 8(d): fixed op counts per radial point, weights in flop (DFMA = 2).
# weights = FP64 flop of the fast path of CUDA 12.9's sm_100a math functions, counted in SASS
W_POW, W_EXP, W_LOG, W_DIV, W_SQRT, W_ADDMUL = 100., 31., 45., 20., 20., 1.


def flops_per_model_step(P, S, n_lambda, n):
    """P Picard iterations, S quadrature trial steps per point of the Lx integral."""
    solve = P * (2 * W_POW + 6 * W_DIV + 10 * W_ADDMUL) + (4 * W_DIV + 25 * W_ADDMUL) + (2 * W_POW + 3 * W_DIV)
    row = (12 * W_POW + 3 * W_LOG + 2 * W_SQRT + 6 * W_DIV + 60 * W_ADDMUL) + S * (6 * W_EXP + 6 * W_DIV + 60 * W_ADDMUL) + \
        n_lambda * (W_EXP + 2 * W_DIV + 8 * W_ADDMUL)
    return (solve + row) * n


def workload_args(n_gpus):
    """The global ensemble: alpha(40 * n_gpus) x Cirr(25), quasistat, Thot=1e4, boundcond=Tirr, isotropic."""
    from freddi_b200.workloads import config2_args
    return config2_args(np.linspace(0.1, 1.0, N_ALPHA * n_gpus), np.geomspace(1e-5, 5e-3, N_CIRR))


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


def dist_setup(n_gpus):
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    dist = None
    if world > 1:
        import torch.distributed as dist
    return rank, world, local, dist


def cpu_baseline(n_models_cap=None, cores=None, args=None, sweep='alpha x Cirr'):
    """The reference's own CPU implementation (oracle/_ref: unmodified sources + Boost shim), one model per
    thread across all host cores, on a bounded sample of the N=1 workload (or of `args`: scripts/sweep_1e5.py)."""
    from oracle import pyoracle
    ref = pyoracle.Ref()
    cores = cores or os.cpu_count() or 1
    if args is None:
        args = workload_args(1)
    n = min(len(args), n_models_cap or max(8, 8 * cores))
    idx = np.linspace(0, len(args) - 1, n).round().astype(int)  # stride-sampled over the sweep
    sample = [args[i] for i in idx]
    out = ref.run_ensemble(sample, LAMBDAS, n_threads=cores)
    value = out['steps'] / out['evolve_s']
    return dict(value=value, unit=UNIT, cores=cores, kind='reference',
                sample='{} of {} models (every {:.1f}-th of the {} sweep), all {} steps each; evolve {:.2f} s, '
                       'model construction {:.2f} s excluded'.format(n, len(args), len(args) / n, sweep, out['steps'] // n,
                                                                      out['evolve_s'], out['construct_s']),
                steps=out['steps'], evolve_s=out['evolve_s'], construct_s=out['construct_s'])


def run_reference(opts):
    rank, world, local, dist = dist_setup(opts.gpus)
    if rank != 0:
        return
    times, steps = [], 0
    last = None
    for k in range(opts.warmup + opts.steps):
        last = cpu_baseline(n_models_cap=opts.ref_models)
        if k >= opts.warmup:
            times.append(last['evolve_s'])
            steps += last['steps']
    value = steps / sum(times)
    line = dict(impl='reference', metric=METRIC, value=value, unit=UNIT, n_gpus=opts.gpus, steps=opts.steps, warmup=opts.warmup,
                ms_per_step=1e3 * sum(times) / len(times), higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f64',
                data='synthetic',
                config=dict(workload='BASELINE configs[1]: BH alpha(40) x Cirr(25) sweep, Nx=1000, 200 steps, 3 band fluxes per row',
                            sample=last['sample']),
                cpu_baseline=dict(value=value, unit=UNIT, cores=last['cores'], kind='reference', sample=last['sample']),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                light_curves_per_s=value / 200.)
    print(json.dumps(line), flush=True)


def run_gpu(opts):
    import torch
    from freddi_b200 import Ensemble
    from freddi_b200.ensemble import measure_fp64_peak, shard_indices
    from freddi_b200.args import args_array, FB200Args
    import ctypes as C

    rank, world, local, dist = dist_setup(opts.gpus)
    torch.cuda.set_device(local)
    if dist is not None:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    n_gpus = world if world > 1 else 1

    all_args = workload_args(n_gpus)
    mine = [all_args[i] for i in shard_indices(len(all_args), rank, n_gpus)]
    arr = args_array(mine)
    n_local = len(mine)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')  # > 126 MB L2

    total_iters = opts.warmup + opts.steps
    ensembles = [Ensemble(arr, lambdas=LAMBDAS, device=local) for _ in range(total_iters)]
    for e in ensembles:
        e.evolve(0)  # row 0 + first-touch of every buffer: timed steps start with inputs resident in HBM
    sampler = ClockSampler(local)
    dev_ms = []
    model_steps = 0
    picard = lx = 0
    for k, e in enumerate(ensembles):
        if k == opts.warmup:
            barrier()
            sampler.start()
            t_wall0 = time.perf_counter()
        flush_buf.fill_(k & 0xff)  # evict L2 between iterations
        torch.cuda.synchronize()
        e.evolve_async(-1)
        e.sync()
        if k >= opts.warmup:
            dev_ms.append(e.last_evolve_ms())
            st, valid, _ = e.status()
            model_steps += int((valid - 1).sum())
            p, l = e.counters()
            picard += int(p.sum())
            lx += int(l.sum())
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    launches = sum(e.last_launches() for e in ensembles[opts.warmup:])
    rows_bytes = ensembles[0].n_models * ensembles[0].n_rows * ensembles[0].n_cols * 8
    n_rows = ensembles[0].n_rows
    for e in ensembles:
        e.close()

    # ---- end to end through the public API with host buffers ----
    # One step = reset() [resolved initial conditions: pinned host -> device] + evolve() + rows() [device -> pinned host].
    # Argument resolution / model construction is outside, as it is for the reference arm (its construction time is
    # reported separately and excluded); the fully inclusive variant is timed too (e2e.with_construction).
    pinned = torch.empty((n_local, n_rows, ensembles[0].n_cols), dtype=torch.float64).pin_memory().numpy()
    e2e_times, e2e_full_times = [], []
    with Ensemble(arr, lambdas=LAMBDAS, device=local) as e:
        for k in range(opts.warmup + opts.steps):
            barrier()
            t0 = time.perf_counter()
            e.reset()
            e.evolve(-1)
            e.rows(out=pinned)
            if k >= opts.warmup:
                e2e_times.append(time.perf_counter() - t0)
        h2d = n_local * (1000 * 8 + 64)
    for k in range(1 + opts.steps):
        barrier()
        t0 = time.perf_counter()
        with Ensemble(arr, lambdas=LAMBDAS, device=local) as e:
            e.evolve(-1)
            e.rows(out=pinned)
        if k >= 1:
            e2e_full_times.append(time.perf_counter() - t0)
    h2d_full = n_local * (C.sizeof(FB200Args) + 2 * 1000 * 8 + 512 + 64)

    # ---- reduce over ranks: time = max, work = sum ----
    t_dev = sum(dev_ms) / 1e3
    t_e2e = sum(e2e_times)
    t_e2e_full = sum(e2e_full_times)
    red = torch.tensor([t_dev, t_wall, t_e2e, t_e2e_full], dtype=torch.float64, device='cuda')
    work = torch.tensor([model_steps, picard, lx, launches], dtype=torch.float64, device='cuda')
    if dist is not None:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
    t_dev, t_wall, t_e2e, t_e2e_full = red.tolist()
    model_steps_all, picard_all, lx_all, launches_all = work.tolist()

    if rank == 0:
        value = model_steps_all / t_dev
        P = picard_all / model_steps_all
        S = lx_all / (model_steps_all + n_gpus * n_local * opts.steps) / 1000.  # per point per row (rows = steps + row 0)
        fl = flops_per_model_step(P, S, len(LAMBDAS), 1000)
        try:
            peak = measure_fp64_peak(local)
        except Exception:
            peak = None
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        traffic = None
        try:  # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu --set full capture
            traffic = json.load(open(os.path.join(ROOT, 'profiles', 'roofline_traffic.json')))['dram_bytes_per_launch']
        except Exception:
            pass
        per_gpu_rate = value / n_gpus
        achieved_tf = per_gpu_rate * fl / 1e12
        hbm_bytes_per_step = 8 * (15 + len(LAMBDAS))
        roofline = dict(bound='fp64', achieved=achieved_tf, peak=peak, unit='TFLOP/s',
                        frac=(achieved_tf / peak) if peak else None, traffic=traffic,
                        peak_source='DFMA micro-kernel measured in this run (MEASURED_PEAKS.json has no FP64 entry)',
                        flop_model='SURVEY.md 8(d) op counts x SASS-counted FP64 flop per call (pow 100, exp 31, log 45, div 20, sqrt 20); '
                                   'measured P={:.2f} Picard iterations/step, S={:.2f} quadrature trial steps/point/row (the trial steps the kernel '
                                   'executes: rings it prunes are not counted)'.format(P, S),
                        flops_per_model_step=fl,
                        hbm=dict(achieved=per_gpu_rate * hbm_bytes_per_step / 1e9, peak=peaks.get('hbm_gbs'), unit='GB/s',
                                 note='algorithmic HBM traffic is one {}-byte row per model-step: irrelevant to this path'.format(
                                     hbm_bytes_per_step)))
        cpu = None
        if n_gpus == 1 and not opts.no_cpu_baseline:
            try:
                cpu = cpu_baseline(n_models_cap=opts.ref_models)
            except Exception as exc:  # the checker library is test infrastructure; the bench line survives without it
                cpu = dict(value=None, unit=UNIT, cores=os.cpu_count(), kind='reference', sample='unavailable: {}'.format(exc))
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=n_gpus, steps=opts.steps, warmup=opts.warmup,
                    ms_per_step=1e3 * t_dev / opts.steps, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f64',
                    data='synthetic',
                    config=dict(workload='BASELINE configs[1]: BH alpha({}) x Cirr({}) sweep = {} models/GPU, Nx=1000, 200 steps, '
                                         '{} band fluxes per row'.format(N_ALPHA * n_gpus, N_CIRR, n_local, len(LAMBDAS)),
                                l2='flushed (256 MB write) before every timed iteration',
                                sharding='interleaved model index, no collective on the data path',
                                wall_ms_per_step=1e3 * t_wall / opts.steps),
                    roofline=roofline, cpu_baseline=cpu, clocks=clocks,
                    e2e=dict(value=model_steps_all / t_e2e, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=rows_bytes,
                             iter_ms=[round(1e3 * x, 2) for x in e2e_times],
                             note='per step: reset() [initial F + state, pinned host -> device] + evolve() + rows() [device -> pinned host], '
                                  'wall clock; model construction excluded as in the reference arm',
                             with_construction=dict(value=model_steps_all / t_e2e_full, unit=UNIT, h2d_bytes_per_step=h2d_full,
                                                    iter_ms=[round(1e3 * x, 2) for x in e2e_full_times],
                                                    note='Ensemble(args) [host argument resolution on all cores + allocation + H2D] + '
                                                         'evolve + rows() + destroy')),
                    gpu_launches=int(launches_all), light_curves_per_s=value / 200.,
                    model_steps_per_iteration=model_steps_all / opts.steps)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--ref-models', type=int, default=None, help='models in the CPU sample (default 8 x cores, max 1000)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    opts = ap.parse_args()
    if opts.impl == 'reference':
        run_reference(opts)
    else:
        run_gpu(opts)


if __name__ == '__main__':
    main()
