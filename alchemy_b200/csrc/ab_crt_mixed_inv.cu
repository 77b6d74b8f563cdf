// ab_crt_mixed_inv.cu -- the inverse instantiation of the mixed-radix transform kernel (see ab_crt_mixed_kernel.cuh)
#define AB_MX_KERNEL_INV 1
#include "ab_crt_mixed_kernel.cuh"
