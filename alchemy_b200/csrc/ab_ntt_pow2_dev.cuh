// ab_ntt_pow2_dev.cuh -- device building blocks of the power-of-two transforms, shared by ab_ntt_pow2.cu (batched CRT / crtInv) and
// ab_mulks_n256.cu (the PT2CT mul_ pipeline of examples/Arithmetic.hs in one kernel): Harvey / exact butterflies, the window-packed
// twiddle access and the register windows (LOGE butterfly stages on 2^LOGE residues held by one thread).  See ab_ntt_pow2.cu for the
// mathematics and the table layout (ab_internal.h: ab_pow2_*).
#pragma once
#include <utility>

#include "ab_internal.h"
#include "ab_modarith.cuh"

struct Pow2Limb {
    uint32_t q, ninv, ninv_s, pad;
    const uint2 *tw;          // window-packed table (forward or inverse)
    uint2 tw0[32];            // window 0, entry k = 2^lb + ip
};
struct Pow2Args {
    Pow2Limb limb[AB_MAX_LIMBS];
    int32_t nlimbs;
    int64_t batch;            // ring elements; transform (b, limb) lives at y + (b * nlimbs + limb) * N
    int32_t pf_dist;          // cluster kernels: CTAs ahead (in launch order) whose input is pulled into L2; 0 = no prefetch
};

// kernels with two CTAs per transform (n = 32768; the split inverse of n = 16384): the half-transform of the CTA `dist` places later in launch
// order (contiguous) is requested into L2 now, 8 KB per thread of the first threads, so that its load phase waits on L2 instead of HBM
template <int LOGN>
__device__ __forceinline__ void pow2_c2_prefetch(const Pow2Args &args, const int64_t *y) {
    if (args.pf_dist <= 0 || threadIdx.x >= (1 << (LOGN - 11))) return;
    const int64_t lin = (int64_t)blockIdx.y * gridDim.x + blockIdx.x + args.pf_dist;
    if (lin >= (int64_t)gridDim.x * gridDim.y) return;
    const int64_t x2 = lin % gridDim.x, l2 = lin / gridDim.x;
    const int64_t *src = y + ((x2 >> 1) * args.nlimbs + l2) * (1 << LOGN) + (x2 & 1) * (1 << (LOGN - 1)) + threadIdx.x * 1024;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(8192) : "memory");
}
// the same for the one-transform-per-CTA kernels (n = 8192, 16384): the transform of the CTA `dist` places later, 8 KB per thread
template <int LOGN>
__device__ __forceinline__ void pow2_prefetch(const Pow2Args &args, const int64_t *y) {
    if (args.pf_dist <= 0 || threadIdx.x >= (1 << LOGN) / 1024) return;
    const int64_t lin = (int64_t)blockIdx.y * gridDim.x + blockIdx.x + args.pf_dist;
    if (lin >= (int64_t)gridDim.x * gridDim.y) return;
    const int64_t b2 = lin % gridDim.x, l2 = lin / gridDim.x;
    const int64_t *src = y + ((b2 * args.nlimbs + l2) << LOGN) + threadIdx.x * 1024;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(8192) : "memory");
}

template <bool LAZY>
__device__ __forceinline__ void bf_fwd(uint32_t &X, uint32_t &Y, uint2 w, uint32_t q) {
    if (LAZY) {   // Harvey: inputs in [0,4q), outputs in [0,4q); needs q < 2^30
        uint32_t x = ab_csub(X, 2 * q);
        uint32_t t = ab_shoup_lazy(Y, w.x, w.y, q);
        X = x + t;
        Y = x - t + 2 * q;
    } else {      // fully reduced, q < 2^31
        uint32_t t = ab_shoup(Y, w.x, w.y, q);
        uint32_t x = X;
        X = ab_addmod(x, t, q);
        Y = ab_submod(x, t, q);
    }
}
template <bool LAZY>
__device__ __forceinline__ void bf_inv(uint32_t &X, uint32_t &Y, uint2 w, uint32_t q) {
    if (LAZY) {   // Gentleman-Sande: inputs and outputs in [0,2q)
        uint32_t s = X + Y, d = X - Y + 2 * q;
        X = ab_csub(s, 2 * q);
        Y = ab_shoup_lazy(d, w.x, w.y, q);
    } else {
        uint32_t s = ab_addmod(X, Y, q), d = ab_submod(X, Y, q);
        X = s;
        Y = ab_shoup(d, w.x, w.y, q);
    }
}

__device__ __forceinline__ int phys(int a) { return a + (a >> 5); }

// compile-time loops: the register array x[] must only ever be indexed by constants (a loop the compiler
// declines to unroll would push it to local memory)
template <class F, int... I>
__device__ __forceinline__ void static_for_impl(F &&f, std::integer_sequence<int, I...>) {
    (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F &&f) {
    static_for_impl(f, std::make_integer_sequence<int, N>{});
}

// twiddle k of this thread's column: window 0 from kernel constants (CTA-uniform), later windows from the
// packed table, two entries per 16-byte load
template <int W, int S, int K>
__device__ __forceinline__ uint2 tw_get(const Pow2Limb &lp, const uint4 *tp, uint4 &pr, bool first_of_stage) {
    if (W == 0) return lp.tw0[K];
    if ((K & 1) == 0 || first_of_stage) pr = __ldg(tp + ((int64_t)(K >> 1) << S));
    return (K & 1) ? make_uint2(pr.z, pr.w) : make_uint2(pr.x, pr.y);
}

template <int LOGE, int W, int S, int LB0, bool LAZY>
__device__ __forceinline__ void fwd_window(uint32_t (&x)[1 << LOGE], const Pow2Limb &lp, const uint2 *__restrict__ tw, int64_t base,
                                           int low, uint32_t q) {
    constexpr int KOFF = (LB0 == 0) ? 0 : (1 << LB0);
    const uint4 *tp = reinterpret_cast<const uint4 *>(tw + base) + low;
    uint4 pr = make_uint4(0, 0, 0, 0);
    static_for<(LOGE > LB0 ? LOGE - LB0 : 0)>([&](auto LBI) {
        constexpr int lb = LB0 + decltype(LBI)::value;
        static_for<(1 << lb)>([&](auto IP) {
            constexpr int ip = decltype(IP)::value, k = (1 << lb) + ip - KOFF;
            const uint2 w = tw_get<W, S, k>(lp, tp, pr, ip == 0);
            static_for<(1 << (LOGE - 1 - lb))>([&](auto HI) {
                constexpr int i = ip + (decltype(HI)::value << (lb + 1));
                bf_fwd<LAZY>(x[i], x[i | (1 << lb)], w, q);
            });
        });
    });
}

template <int LOGE, int W, int S, int LB0, bool LAZY>
__device__ __forceinline__ void inv_window(uint32_t (&x)[1 << LOGE], const Pow2Limb &lp, const uint2 *__restrict__ tw, int64_t base,
                                           int low, uint32_t q) {
    constexpr int KOFF = (LB0 == 0) ? 0 : (1 << LB0);
    const uint4 *tp = reinterpret_cast<const uint4 *>(tw + base) + low;
    uint4 pr = make_uint4(0, 0, 0, 0);
    static_for<(LOGE > LB0 ? LOGE - LB0 : 0)>([&](auto LBI) {
        constexpr int lb = LOGE - 1 - decltype(LBI)::value;
        static_for<(1 << lb)>([&](auto IP) {
            constexpr int ip = decltype(IP)::value, k = (1 << lb) + ip - KOFF;
            const uint2 w = tw_get<W, S, k>(lp, tp, pr, ip == 0);
            static_for<(1 << (LOGE - 1 - lb))>([&](auto HI) {
                constexpr int i = ip + (decltype(HI)::value << (lb + 1));
                if (S + lb == 0) {
                    // last stage of the inverse: fold n^-1 in (tw[1] already carries it for the lower output)
                    const uint32_t X = x[i], Y = x[i | 1];
                    if (LAZY) {
                        x[i] = ab_shoup_lazy(X + Y, lp.ninv, lp.ninv_s, q);
                        x[i | 1] = ab_shoup_lazy(X - Y + 2 * q, w.x, w.y, q);
                    } else {
                        x[i] = ab_shoup(ab_addmod(X, Y, q), lp.ninv, lp.ninv_s, q);
                        x[i | 1] = ab_shoup(ab_submod(X, Y, q), w.x, w.y, q);
                    }
                } else {
                    bf_inv<LAZY>(x[i], x[i | (1 << lb)], w, q);
                }
            });
        });
    });
}
