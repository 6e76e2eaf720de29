#!/usr/bin/env python
"""Where the end-to-end time of one ensemble goes: create (host resolve + allocation + H2D), evolve, rows D2H."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from freddi_b200 import Ensemble
from freddi_b200.args import args_array
from freddi_b200.workloads import config2_args, LAM3
args = config2_args(np.linspace(0.1, 1.0, 40), np.geomspace(1e-5, 5e-3, 25))
arr = args_array(args)
out = None
for it in range(6):
    t0 = time.perf_counter()
    e = Ensemble(arr, lambdas=LAM3)
    t1 = time.perf_counter()
    e.evolve(-1)
    t2 = time.perf_counter()
    if out is None:
        out = np.empty((e.n_models, e.n_rows, e.n_cols))
    e.rows(out=out)
    t3 = time.perf_counter()
    e.close()
    t4 = time.perf_counter()
    print('iter %d: create %.1f ms, evolve %.1f ms (device %.1f), rows %.1f ms, destroy %.1f ms, total %.1f ms' % (
        it, 1e3 * (t1 - t0), 1e3 * (t2 - t1), e_ms if (e_ms := 0) else 0, 1e3 * (t3 - t2), 1e3 * (t4 - t3), 1e3 * (t4 - t0)))
