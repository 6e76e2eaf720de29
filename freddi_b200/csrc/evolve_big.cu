// Large-grid build of the evolve kernel (1024 < Nx <= 16384): evolve.cu compiled a second time with the per-point
// arrays in the CTA's global scratch slot instead of shared memory (gpu_exec.cuh, FB_BIG).  The reference's analytical
// tests need it (python/test/test_analytical.py:84,113,216 run Nx = 10^4); the BASELINE configurations do not.
#define FB_BIG 1
#define FB_VARIANT big
#define FB200_NX_MAX 16384
#define fb200 fb200_big
#include "evolve.cu"
