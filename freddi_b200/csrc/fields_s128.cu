// Small-grid build of the read-out kernels, see evolve_s128.cu.
#define FB_VARIANT s128
#define FB200_NX_MAX 128
#define FB_NT 32
#define FB_CTAS_PER_SM 12
#define fb200 fb200_s128
#include "fields.cu"
