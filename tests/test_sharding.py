"""The N>1 path on CPU: world_size-2 gloo.  Each rank takes its interleaved shard of a small sweep, evolves it (with
the host-emulated engine — no GPU here), rank 0 gathers the rows and they must equal the single-process result;
the timing reduction takes the max of the times and the sum of the work."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import common as cm
from freddi_b200.sharding import shard_indices, gather_rows, reduce_timing


def _sweep():
    return cm.config2_args([0.2, 0.5, 0.8], [1e-4, 1e-3], time=2.0, Nx=64)


def _evolve(emu, args_list):
    rows = []
    for a in args_list:
        rows.append(emu.run(a, 8, 64, cm.LAM3)['rows'])
    return np.stack(rows)


def _worker(rank, world, port, out_path):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    emu = cm.HostEmu()
    args = _sweep()
    idx = shard_indices(len(args), rank, world)
    local = _evolve(emu, [args[i] for i in idx])
    full = gather_rows(dist, local, len(args), rank, world)
    (tmax, tsum_check), (steps,) = reduce_timing(dist, torch, [1.0 + rank, 5.0 - rank], [local.shape[0] * 8])
    if rank == 0:
        np.savez(out_path, rows=full, tmax=tmax, t2=tsum_check, steps=steps)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_indices_partition():
    for n, w in ((1000, 8), (10, 4), (7, 2), (3, 8)):
        allidx = np.concatenate([shard_indices(n, r, w) for r in range(w)])
        assert sorted(allidx.tolist()) == list(range(n))
        sizes = [len(shard_indices(n, r, w)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_ensemble_matches_single_process(built, tmp_path):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    out = str(tmp_path / 'gathered.npz')
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    emu = cm.HostEmu()
    expect = _evolve(emu, _sweep())
    np.testing.assert_array_equal(got['rows'], expect)
    assert got['tmax'] == 2.0 and got['t2'] == 5.0  # max over ranks
    assert got['steps'] == 6 * 8                     # sum over ranks
