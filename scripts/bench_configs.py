#!/usr/bin/env python
"""Device-time throughput of the other BASELINE configurations (10^3-model slices) — for the record, not the headline."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from freddi_b200 import Ensemble
from freddi_b200.workloads import config2_args, config3_args, config4_args, config5_args, ini_args, LAM3
cases = {
    'config1 default freddi.ini x1000 (identical models)': ([ini_args()] * 1000, ()),
    'config2 BH alpha(40) x Cirr(25), irradiation, Thot=1e4': (config2_args(np.linspace(0.1, 1.0, 40), np.geomspace(1e-5, 5e-3, 25)), LAM3),
    'config3 NS Bx(25) x alpha(40), Bx ~ 1e8': (config3_args(np.geomspace(10 ** 7.5, 10 ** 8.5, 25), np.linspace(0.1, 1.0, 40)), LAM3),
    'config4 BH Woods1996 wind + irradiation, alpha(40) x Cirr(25), 3 bands': (config4_args(np.linspace(0.1, 1.0, 40), np.geomspace(1e-5, 5e-3, 25)), LAM3),
    'config5 slice Mx(5) x alpha(5) x Mdisk0(5) x Cirr(8)': (config5_args(np.linspace(3, 15, 5), np.linspace(0.1, 1.0, 5), np.geomspace(1e25, 1e27, 5), np.geomspace(1e-5, 5e-3, 8)), LAM3),
}
for name, (args, lam) in cases.items():
    with Ensemble(args, lambdas=lam) as e:
        e.evolve(0)
        best = None
        for rep in range(3):
            e.reset(); e.evolve(0); e.evolve(-1)
            ms = e.last_evolve_ms()
            best = ms if best is None else min(best, ms)
        st, valid, pic = e.status()
        pi, lx = e.counters()
        steps = int((valid - 1).sum())
        print(json.dumps(dict(case=name, models=len(args), model_steps=steps, ms=round(best, 2), model_steps_per_s=round(steps / best * 1e3),
                              collapsed=int((st == 1).sum()), picard_per_step=round(float(pi.sum()) / max(steps, 1), 2),
                              lx_trials_per_ring_row=round(float(lx.sum()) / max(int(valid.sum()), 1) / 1000., 2))))
