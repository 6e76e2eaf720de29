import ctypes as C, sys, os
sys.path.insert(0, '/root/repo')
import numpy as np
from freddi_b200 import Ensemble, _lib
from freddi_b200.workloads import config2_args, LAM3
args = config2_args(np.linspace(0.1, 1.0, 40), np.geomspace(1e-5, 5e-3, 25))
for lx, lam in ((0, ()), (0, LAM3[:1]), (0, LAM3), (1, ())):
    with Ensemble(args, lambdas=lam, compute_lx=bool(lx)) as e:
        e.evolve(0); e.evolve(-1)
        ms = e.last_evolve_ms()
        st, valid, pic = e.status()
        steps = int((valid - 1).sum())
        out = (C.c_longlong * 16)()
        _lib.lib.fb200_debug_phase_cycles.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
        _lib.lib.fb200_debug_phase_cycles(e._h, out)
        print('lx=%d nlam=%d: %.1f ms; phase4 (Lx+Fnu sums) %.0f cycles/model-step; total %.0f' % (lx, len(lam), ms, out[4] / steps, sum(out) / steps))
