#!/usr/bin/env python
"""Per-phase cycle breakdown of the evolve kernel (FB_PROFILE build given through FB200_LIB)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from freddi_b200 import Ensemble, _lib
from freddi_b200.workloads import config2_args, LAM3
args = config2_args(np.linspace(0.1, 1.0, 40), np.geomspace(1e-5, 5e-3, 25))
names = ['0 step setup', '1 PCR shuffle levels + z', '2 W,K update+norm', '3 structure+truncate', '4 Lx+Fnu sums', '5 row write', '6 step head', '7 model load/init/store',
         '8 block elimination', '9 PCR shared-memory levels', '10 TphX pass']
for lx in (1, 0):
    with Ensemble(args, lambdas=LAM3, compute_lx=bool(lx)) as e:
        e.evolve(0); e.evolve(-1)
        ms = e.last_evolve_ms()
        st, valid, pic = e.status()
        steps = int((valid - 1).sum())
        out = (C.c_longlong * 16)()
        _lib.lib.fb200_debug_phase_cycles.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
        _lib.lib.fb200_debug_phase_cycles(e._h, out)
        tot = sum(out)
        print('compute_lx=%d: %.1f ms, %d model-steps, %.0f steps/s, %.0f cycles/model-step (thread-0 clock)' % (lx, ms, steps, steps / ms * 1e3, tot / max(steps, 1)))
        for n, c in zip(names, out):
            print('   %-26s %6.2f%%  %9.0f cycles/model-step' % (n, 100. * c / max(tot, 1), c / max(steps, 1)))
