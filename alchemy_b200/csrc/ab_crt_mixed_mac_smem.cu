// ab_crt_mixed_mac_smem.cu -- the fused gadget inner product with its running sums in shared memory (see ab_crt_mixed_mac_kernel.cuh)
#define AB_MX_MAC_SMEM 1
#include "ab_crt_mixed_mac_kernel.cuh"
