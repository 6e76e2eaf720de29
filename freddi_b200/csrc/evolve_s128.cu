// Small-grid build of the evolve kernel: one WARP per model (Nx <= 128): 32 threads x 4 rings, ~17 KB of shared memory per model, 12 models resident per SM.
// Same engine, same kernel source (evolve.cu) with a smaller CTA: at small Nx the 256-thread layout of the main build leaves
// most of a CTA idle in every per-ring pass (240 of 256 threads at Nx = 64), and its 114 KB of shared memory per model
// caps an SM at two models.  (BASELINE.json north_star: "one CTA (or warp, for small Nx) owns a model".)
#define FB_VARIANT s128
#define FB200_NX_MAX 128
#define FB_NT 32
#define FB_CTAS_PER_SM 12
#define fb200 fb200_s128
#include "evolve.cu"
