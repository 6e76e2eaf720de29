// Mid-size build of the read-out kernels, see evolve_s512.cu.
#define FB_VARIANT s512
#define FB200_NX_MAX 512
#define FB_NT 128
#define FB_CTAS_PER_SM 3
#define fb200 fb200_s512
#include "fields.cu"
