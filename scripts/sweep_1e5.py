#!/usr/bin/env python
"""BASELINE configs[4] at full size: the 10^5-model Mx(10) x alpha(20) x Mdisk0(25) x Cirr(20) sweep, Nx = 1000,
200 steps, 3 band fluxes per row, sharded over the GPUs of one box by interleaved model index (no collective on the data path), next
to the reference's CPU implementation (bench.py's cpu_baseline leg) on a stride sample of the same sweep on all host cores.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 scripts/sweep_1e5.py
    python scripts/sweep_1e5.py --models 12500          # one GPU, an eighth of the sweep

One JSON line on rank 0.  Device time: CUDA events around evolve(), max over ranks; end to end: Ensemble(args) [host argument
resolution + allocation + H2D] + evolve() + rows() into a pinned host buffer, wall clock between barriers, max over ranks."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

LAMBDAS = [3e-5, 5e-5, 8e-5]  # cm: 3000, 5000, 8000 A


def sweep_args(n_models):
    from freddi_b200.workloads import config5_args
    args = config5_args(np.linspace(3, 15, 10), np.linspace(0.1, 1.0, 20), np.geomspace(1e25, 1e27, 25), np.geomspace(1e-5, 5e-3, 20))
    if n_models and n_models < len(args):
        idx = np.linspace(0, len(args) - 1, n_models).round().astype(int)
        args = [args[i] for i in idx]
    return args


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--models', type=int, default=0, help='stride-sampled subset of the 10^5 sweep (0: all)')
    ap.add_argument('--ref-models', type=int, default=0, help='models of the CPU sample (default 8 per host core)')
    ap.add_argument('--no-cpu', action='store_true')
    opts = ap.parse_args()
    import torch
    from freddi_b200 import Ensemble
    from freddi_b200.args import args_array
    from freddi_b200.ensemble import shard_indices
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    torch.cuda.set_device(local)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    args = sweep_args(opts.models)
    n_total = len(args)
    mine = [args[i] for i in shard_indices(n_total, rank, world)]
    arr = args_array(mine)
    # warm-up: context, module load, first-touch
    with Ensemble(arr[:min(len(mine), 296)], lambdas=LAMBDAS, device=local) as e:
        e.evolve(-1)
    barrier()
    t0 = time.perf_counter()
    ens = Ensemble(arr, lambdas=LAMBDAS, device=local)
    t1 = time.perf_counter()
    ens.evolve(-1)
    dev_ms = ens.last_evolve_ms()
    t2 = time.perf_counter()
    out = torch.empty((ens.n_models, ens.n_rows, ens.n_cols), dtype=torch.float64).pin_memory().numpy()
    t3 = time.perf_counter()
    ens.rows(out=out)
    t4 = time.perf_counter()
    st, valid, pic = ens.status()
    steps = int((valid - 1).sum())
    collapsed = int((st == 1).sum())
    finite = bool(np.isfinite(out[np.arange(len(mine)), 0, 1]).all())  # Mdot of row 0 of every model
    ens.close()
    barrier()
    e2e_s = (t2 - t0) + (t4 - t3)  # the pinned result buffer is allocated once per process, not per sweep
    red = torch.tensor([dev_ms / 1e3, e2e_s, t1 - t0, t4 - t3], dtype=torch.float64, device='cuda')
    tot = torch.tensor([steps, collapsed, len(mine), int(pic.sum())], dtype=torch.float64, device='cuda')
    if dist is not None:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    t_dev, t_e2e, t_create, t_rows = red.tolist()
    steps_all, collapsed_all, models_all, picard_all = tot.tolist()
    if rank == 0:
        line = dict(metric='ensemble disc-model steps/sec at Nx=1000', unit='model-steps/s', n_gpus=world,
                    config=dict(workload='BASELINE configs[4]: Mx(10) x alpha(20) x Mdisk0(25) x Cirr(20) = {} models, Nx=1000, 200 steps, '
                                         '3 band fluxes per row'.format(int(models_all)),
                                sharding='interleaved model index, {} models per GPU, no collective on the data path'.format(len(mine))),
                    value=steps_all / t_dev, evolve_s=t_dev, model_steps=steps_all, collapsed_models=int(collapsed_all),
                    picard_per_step=picard_all / steps_all, light_curves_per_s=models_all / t_dev,
                    e2e=dict(value=steps_all / t_e2e, seconds=t_e2e, create_s=t_create, rows_d2h_s=t_rows,
                             light_curves_per_s=models_all / t_e2e,
                             note='Ensemble(args) [host argument resolution + allocation + H2D] + evolve + rows() into pinned host memory, '
                                  'wall clock, max over ranks'),
                    row0_finite=finite)
        if not opts.no_cpu:
            import bench  # its cpu_baseline leg: the reference's CPU implementation on a stride sample, all host cores
            cores = os.cpu_count() or 1
            cpu = bench.cpu_baseline(n_models_cap=opts.ref_models or 8 * cores, cores=cores, args=sweep_args(0),
                                     sweep='Mx x alpha x Mdisk0 x Cirr')
            cpu['extrapolated_full_sweep_s'] = steps_all / cpu['value']
            line['cpu_baseline'] = cpu
            line['speedup_vs_all_core_host'] = dict(device_timed=line['value'] / cpu['value'], end_to_end=line['e2e']['value'] / cpu['value'])
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
