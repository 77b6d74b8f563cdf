// ab_kernels.h -- launchers of the streaming kernels in ab_elementwise.cu (internal).
#pragma once
#include "ab_internal.h"

void ab_launch_index_map(int mode, int64_t *out, const int64_t *in, const int32_t *idx, const int8_t *sgn, const ab_plan *p,
                         int64_t rows, int64_t n_out, int64_t n_in, int64_t rel, int32_t nlimbs, cudaStream_t st);
void ab_launch_twace_crt(const ab_ext *e, int64_t *out, const int64_t *in, int64_t batch, int32_t nlimbs, cudaStream_t st);
int32_t ab_gadget_digits(int64_t q, int32_t base);
void ab_launch_decompose(const ab_plan *p, int64_t *out, const int64_t *in, int64_t batch, int32_t nlimbs, int32_t out_limbs,
                         int32_t base, cudaStream_t st);
void ab_launch_rescale_drop(const ab_plan *p, int64_t *out, const int64_t *in, int64_t batch, int32_t nlimbs, int32_t drop, cudaStream_t st);
void ab_launch_lift_reduce(const ab_plan *p, int64_t *out, int64_t out_bstride, const int64_t *in, int64_t in_bstride, int64_t batch,
                           int32_t src_limb, const int32_t *dst_limbs, int32_t nout, cudaStream_t st);
void ab_launch_sub_scale(const ab_plan *p, int64_t *out, int64_t out_bstride, const int64_t *in, int64_t in_bstride,
                         const int64_t *corr, int64_t corr_bstride, int64_t batch, int32_t nl, const int64_t *scal, cudaStream_t st);
void ab_launch_ct_mul(const ab_plan *p, int64_t *out3, const int64_t *a2, const int64_t *b2, int64_t batch, int32_t nlimbs, cudaStream_t st);
// prescale (HOST, nlimbs residues, may be null): every coefficient of the product is also multiplied by prescale[l] (toMSD: p^-1 mod q_l)
void ab_launch_ct_mul_up(const ab_plan *p, int64_t *out3, const int64_t *a2, const int64_t *b2, int64_t batch, int32_t nlimbs, int32_t nout, cudaStream_t st,
                         const int64_t *prescale = nullptr);
void ab_launch_hint_ip(const ab_plan *p, int64_t *out, int64_t out_bstride, const int64_t *dig, int64_t dig_bstride, const int64_t *hint,
                       const int64_t *add, int64_t add_bstride, bool accumulate, int64_t batch, int32_t nlimbs, int32_t ndig, cudaStream_t st);
// hints pre-multiplied by 2^32 / 2^48 (uint2 per residue), then the inner product with one reduction per output
void ab_launch_hint_prep(const ab_plan *p, void *out, const int64_t *hint, int64_t rows, int32_t nlimbs, cudaStream_t st);
// [ndig][2][L][n] int64 -> [ndig][L][n] uint2 (both components x 2^32) for k_crt_mixed_mac
void ab_launch_hint_prep2(const ab_plan *p, void *out, const int64_t *hint, int64_t ndig, int32_t nlimbs, cudaStream_t st);
void ab_launch_hint_ip_m(const ab_plan *p, int64_t *out, int64_t out_bstride, const int64_t *dig, int64_t dig_bstride, const void *hm,
                         const int64_t *add, int64_t add_bstride, int64_t batch, int32_t nlimbs, int32_t ndig, cudaStream_t st);
void ab_launch_lin_ip(const ab_plan *p, int64_t *out, int64_t out_bstride, const int64_t *x, int64_t x_bstride, const int64_t *f,
                      int64_t batch, int32_t nlimbs, int32_t nk, cudaStream_t st);
// the same with x in a sub-ring (n_in rows per limb), replicated through idx[j] (embedCRT)
void ab_launch_lin_ip_rep(const ab_plan *p, int64_t *out, int64_t out_bstride, const int64_t *x, int64_t x_bstride, const int64_t *f,
                          const int32_t *idx, int64_t n_in, int64_t batch, int32_t nlimbs, int32_t nk, cudaStream_t st);
void ab_launch_copy_rows(int64_t *out, int64_t ob, const int64_t *in, int64_t ib, int64_t len, int64_t batch, cudaStream_t st);
void ab_launch_rescale_up(const ab_plan *p, int64_t *out, int64_t ob, const int64_t *in, int64_t ib, int64_t batch, int32_t nfrom, int32_t nto, cudaStream_t st);

// out[b][k][l][j] = sgn[k][j] * in[b*in_bstride + l*n_in + tab[k][j]]  (tab < 0: zero); coeffs composed with embed
void ab_launch_tunnel_gather(const ab_plan *ps, int64_t *out, const int64_t *in, int64_t in_bstride, int64_t n_in, const int32_t *tab,
                             const int8_t *sgn, int64_t batch, int64_t K, int32_t nlimbs, cudaStream_t st);
// out[b*ob + r] += in[b*ib + r]  (r over nlimbs*n)
void ab_launch_add_rows(const ab_plan *p, int64_t *out, int64_t ob, const int64_t *in, int64_t ib, int64_t batch, int32_t nlimbs, cudaStream_t st);

// strided-batch CRT used by the fused ops: `rows` independent n-vectors of ONE limb, row r at y + r*row_stride
void ab_crt_rows(const ab_plan *p, int64_t *y, int64_t rows, int32_t limb, bool inverse, cudaStream_t st);
// generic dispatcher (pow2 fast path or mixed-radix program)
void ab_crt_dispatch(const ab_plan *p, const AbView &v, bool inverse, cudaStream_t st);
