#!/usr/bin/env python
"""Throughput of the small-grid kernel builds (one / two warps per model) against the 256-thread main build at small Nx.

    python scripts/small_nx_bench.py            # the build capi picks (s128 for Nx <= 128, s256 for Nx <= 256)
    FB200_BUILD=main python scripts/small_nx_bench.py   # everything through the main build (the layout of round 1)

Prints one JSON line per Nx: model-steps/s device-timed (CUDA events around the launch), models, ms."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from freddi_b200 import Ensemble  # noqa: E402
from freddi_b200.workloads import config2_args, LAM3  # noqa: E402


def main():
    n_models = int(os.environ.get('N_MODELS', '4000'))
    for Nx in (64, 128, 192, 256, 512, 1000):
        args = config2_args(np.linspace(0.1, 1.0, n_models // 25), np.geomspace(1e-5, 5e-3, 25), Nx=Nx)
        best = None
        with Ensemble(args, lambdas=LAM3) as ens:
            for _ in range(3):
                ens.reset()
                ens.evolve(0)
                ens.evolve(-1)
                ms = ens.last_evolve_ms()
                best = ms if best is None else min(best, ms)
            _, valid, _ = ens.status()
        steps = int((valid - 1).sum())
        print(json.dumps(dict(Nx=Nx, build=os.environ.get('FB200_BUILD', 'auto'), models=len(args), model_steps=steps, ms=best,
                              model_steps_per_s=steps / best * 1e3)), flush=True)


if __name__ == '__main__':
    main()
