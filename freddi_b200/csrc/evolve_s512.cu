// Mid-size build of the evolve kernel: four warps per model (256 < Nx <= 512): 128 threads x 4 rings, ~58 KB of shared memory per
// model, 3 models resident per SM — see evolve_s128.cu.
#define FB_VARIANT s512
#define FB200_NX_MAX 512
#define FB_NT 128
#define FB_CTAS_PER_SM 3
#define fb200 fb200_s512
#include "evolve.cu"
