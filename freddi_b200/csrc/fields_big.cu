// Large-grid build of the read-out kernels, see evolve_big.cu.
#define FB_BIG 1
#define FB_VARIANT big
#define FB200_NX_MAX 16384
#define fb200 fb200_big
#include "fields.cu"
