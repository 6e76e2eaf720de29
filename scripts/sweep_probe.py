import numpy as np, time, sys
sys.path.insert(0, '/root/repo')
import bench
from freddi_b200 import sweep, pinned_empty
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
arr = bench.workload('sweep_1e4_wind', 2, np.arange(0, 2*n, 2))
out = pinned_empty((n,201,18))
for k in range(3):
    t0=time.perf_counter(); r = sweep(arr, lambdas=bench.LAMBDAS, devices=[0], out=out); print("sweep", (time.perf_counter()-t0)*1e3, r.stats, flush=True)
import torch
out2 = torch.empty((n,201,18), dtype=torch.float64).pin_memory().numpy()
for k in range(2):
    t0=time.perf_counter(); r = sweep(arr, lambdas=bench.LAMBDAS, devices=[0], out=out2); print("sweep torch-pinned", (time.perf_counter()-t0)*1e3, r.stats, flush=True)
