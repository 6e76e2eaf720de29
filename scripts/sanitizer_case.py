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
# extras instantiation: passband + companion star columns, batched read-out, star flux at a phase; large-grid build
from freddi_b200.ensemble import Passband
pb = Passband('box', np.linspace(4e-5, 6e-5, 40), np.linspace(4e-5, 6e-5, 40) * 1.0)
args = config2_args([0.3, 0.9], [3e-4], time=1.0, Topt=4000, inclination=40, h2rcold=0.03)
with Ensemble(args, lambdas=LAM3[:2], cold_disk=True, record_F=True, passbands=[pb], star_flux=True) as e:
    e.evolve(-1)
    rows = e.rows()
    e.field_rows('Sigma'); e.flux_rows([5e-5], 'hot'); e.flux_rows([5e-5], 'star', phases=0.5)
    print('extras', rows.shape, 'status', e.status()[0].tolist())
with Ensemble(config2_args([0.5], [1e-4], time=0.5, Nx=1500), lambdas=LAM3[:1], record_F=True) as e:
    e.evolve(-1)
    e.field('Tph', 1); e.mdot_wind()
    print('large grid', e.rows().shape, 'status', e.status()[0].tolist())
print('done')
# round 2: the small-grid builds (one / two / four warps per model), the streamed sweep over two "devices" with tiny scheduling
# groups (FB200_GROUP0 / FB200_GROUP in the environment), the sharded handle, checkpoint / state I/O
from freddi_b200 import sweep
for Nx in (64, 200, 400):
    with Ensemble(config2_args([0.3, 0.9], [1e-4], time=1.0, Nx=Nx), lambdas=LAM3, record_F=True) as e:
        e.evolve(-1)
        e.field('Tph', 1); e.flux([5e-5], 'hot', 2)
        print('Nx', Nx, e.rows().shape, 'status', e.status()[0].tolist())
args = config2_args(np.linspace(0.2, 0.9, 8), np.geomspace(1e-5, 1e-3, 5), time=0.75)
res = sweep(args, lambdas=LAM3, devices=[0, 0])
print('sweep', res.rows.shape, res.stats['launches'], 'launches, status', sorted(set(res.status.tolist())))
with Ensemble(args[:9], lambdas=LAM3, device=[0, 0]) as e:
    e.evolve(1); e.evolve(-1); rows = e.rows(); e.reset(); e.evolve(-1)
    print('sharded', rows.shape, bool((rows.view(np.uint64) == e.rows().view(np.uint64)).all()))
with Ensemble(args[:5], lambdas=LAM3) as a, Ensemble(args[:5], lambdas=LAM3) as b:
    a.evolve(1); b.restore(a.checkpoint()); b.evolve(-1); a.evolve(-1)
    st = a.get_state(); b.set_state(**st)
    print('checkpoint', bool((a.rows().view(np.uint64) == b.rows().view(np.uint64)).all()))
print('done (round 2)')
