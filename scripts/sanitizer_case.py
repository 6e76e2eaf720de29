#!/usr/bin/env python
"""Small all-features case for compute-sanitizer (racecheck / memcheck): BH irradiated + NS + wind ensembles, fields, flux."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from freddi_b200 import Ensemble
from freddi_b200.workloads import config2_args, config3_args, config4_args, LAM3
for args, cold in ((config2_args([0.3, 0.9], [1e-4], time=1.0), True), (config3_args([1e8], [0.25, 0.6], time=1.0), False),
                   (config4_args([0.5], [1e-3, 2e-3], time=1.0), False)):
    with Ensemble(args, lambdas=LAM3, cold_disk=cold, record_F=True) as e:
        e.evolve(2); e.evolve(-1)
        rows = e.rows()
        e.field('Tph', 2); e.flux([5e-5, 8e-5], 'hot', 3); e.mdot_wind()
        e.reset(); e.evolve(1)
        print(len(args), 'models', rows.shape, 'status', e.status()[0].tolist())
print('done')
