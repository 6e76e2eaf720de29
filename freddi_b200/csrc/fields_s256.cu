// Small-grid build of the read-out kernels, see evolve_s256.cu.
#define FB_VARIANT s256
#define FB200_NX_MAX 256
#define FB_NT 64
#define FB_CTAS_PER_SM 7
#define fb200 fb200_s256
#include "fields.cu"
