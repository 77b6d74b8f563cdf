// ab_ntt_pow2.cu -- batched CRT / crtInv for power-of-two cyclotomic index m = 2n (sm_100a).
//
// Replaces lol-cpp `tensorCRTRq` / `tensorCRTInvRq` [RECALL] for m = 2^k, the transform every
// ciphertext operation of Eval.hs:65-67,130,133 funnels into.
//
// Mathematics (Lol conventions, ab_plan.cpp): y[i] = sum_a x[a] psi^(bitrev(a) (2i+1)), psi = omega_2n.
// This is a decimation-in-time network: stage s pairs (a, a + 2^s) and multiplies the upper input
// by W_s[a mod 2^s] = psi^((2l+1) n / 2^(s+1)); input order is bit-reversed, output natural.
//
// Kernel shape: one transform = n/E threads, each holding E = 2^LOGE residues in registers.
// The log2(n) stages are cut into windows of LOGE stages; inside a window every butterfly is
// register-to-register, between windows the residues are transposed through shared memory
// (uint32, padded 33/32 against bank conflicts), in place.  With E = 32 the first two windows
// touch only the 1024-element block owned by one warp, so they need __syncwarp only; a transform
// of n <= 1024 never executes a CTA barrier and n = 32768 executes exactly one.
// HBM traffic is one 16-byte-vector read and one coalesced write of the int64 residues: 16 n
// bytes per transform, the algorithmic minimum.  Twiddles (Shoup pairs) come from L1/L2.
#include "ab_internal.h"
#include "ab_modarith.cuh"

struct Pow2Limb {
    uint32_t q, ninv, ninv_s, pad;
    const uint2 *tw;
};
struct Pow2Args {
    Pow2Limb limb[AB_MAX_LIMBS];
    int32_t nlimbs;
    int64_t ntrans;
};

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

template <int LOGN, int LOGE>
struct Pow2Cfg {
    static constexpr int N = 1 << LOGN, E = 1 << LOGE, T = N / E;
    static constexpr int TPC = (T >= 128) ? 1 : 128 / T;
    static constexpr int THREADS = T * TPC;
    static constexpr int PADN = N + (N >> 5);
    static constexpr int NW = (LOGN + LOGE - 1) / LOGE;
    static constexpr int G = (T < 32) ? T : 32;          // lanes of one warp working on one transform
    static constexpr int MINB = (1024 / THREADS) > 0 ? (1024 / THREADS) : 1;
    __host__ __device__ static constexpr int win_start(int w) { return (w * LOGE + LOGE <= LOGN) ? w * LOGE : LOGN - LOGE; }
    __host__ __device__ static constexpr int win_lb0(int w) { return w * LOGE - win_start(w); }
};

__device__ __forceinline__ int phys(int a) { return a + (a >> 5); }

template <int LOGE, int S, int LB0, bool LAZY>
__device__ __forceinline__ void fwd_window(uint32_t (&x)[1 << LOGE], const uint2 *__restrict__ tw, int low, uint32_t q) {
#pragma unroll
    for (int lb = LB0; lb < LOGE; ++lb) {
        const uint2 *t = tw + (1 << (S + lb)) + low;
#pragma unroll
        for (int ip = 0; ip < (1 << lb); ++ip) {
            const uint2 w = __ldg(t + (ip << S));
#pragma unroll
            for (int hi = 0; hi < (1 << (LOGE - 1 - lb)); ++hi) {
                const int i = ip + (hi << (lb + 1));
                bf_fwd<LAZY>(x[i], x[i | (1 << lb)], w, q);
            }
        }
    }
}

template <int LOGE, int S, int LB0, bool LAZY>
__device__ __forceinline__ void inv_window(uint32_t (&x)[1 << LOGE], const uint2 *__restrict__ tw, int low, uint32_t q,
                                           uint32_t ninv, uint32_t ninv_s) {
#pragma unroll
    for (int lb = LOGE - 1; lb >= LB0; --lb) {
        const uint2 *t = tw + (1 << (S + lb)) + low;
#pragma unroll
        for (int ip = 0; ip < (1 << lb); ++ip) {
            const uint2 w = __ldg(t + (ip << S));
#pragma unroll
            for (int hi = 0; hi < (1 << (LOGE - 1 - lb)); ++hi) {
                const int i = ip + (hi << (lb + 1));
                if (S + lb == 0) {
                    // last stage of the inverse: fold n^-1 in (tw[1] already carries it for the lower output)
                    uint32_t X = x[i], Y = x[i | 1];
                    if (LAZY) {
                        x[i] = ab_shoup_lazy(X + Y, ninv, ninv_s, q);
                        x[i | 1] = ab_shoup_lazy(X - Y + 2 * q, w.x, w.y, q);
                    } else {
                        x[i] = ab_shoup(ab_addmod(X, Y, q), ninv, ninv_s, q);
                        x[i | 1] = ab_shoup(ab_submod(X, Y, q), w.x, w.y, q);
                    }
                } else {
                    bf_inv<LAZY>(x[i], x[i | (1 << lb)], w, q);
                }
            }
        }
    }
}

template <int LOGN, int LOGE, bool LAZY>
__global__ void __launch_bounds__(Pow2Cfg<LOGN, LOGE>::THREADS, Pow2Cfg<LOGN, LOGE>::MINB)
k_crt_pow2_fwd(const Pow2Args args, int64_t *__restrict__ y) {
    using C = Pow2Cfg<LOGN, LOGE>;
    constexpr int E = C::E, T = C::T, G = C::G;
    extern __shared__ uint32_t smem[];
    const int tid = threadIdx.x, tl = tid & (T - 1), slot = tid / T;
    int64_t trans = (int64_t)blockIdx.x * C::TPC + slot;
    const bool active = trans < args.ntrans;
    if (!active) trans = 0;
    const Pow2Limb lp = args.limb[trans % args.nlimbs];
    const uint32_t q = lp.q;
    const uint2 *__restrict__ tw = lp.tw;
    uint32_t *sm = smem + slot * C::PADN;
    int64_t *yp = y + trans * C::N;

    // coalesced 16-byte loads of the block this warp owns in window 0; narrow to uint32 in smem
    {
        const int wb = tl / G, lg = tl % G;
        const int4 *src = reinterpret_cast<const int4 *>(yp + (int64_t)wb * G * E);
        constexpr int CH = (E / 2 >= 8) ? 8 : E / 2;
#pragma unroll
        for (int j0 = 0; j0 < E / 2; j0 += CH) {
            int4 v[CH];
#pragma unroll
            for (int j = 0; j < CH; ++j) v[j] = active ? __ldcs(src + lg + (j0 + j) * G) : make_int4(0, 0, 0, 0);
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                const int a = wb * G * E + 2 * (lg + (j0 + j) * G);
                sm[phys(a)] = (uint32_t)v[j].x;
                sm[phys(a + 1)] = (uint32_t)v[j].z;
            }
        }
    }
    __syncwarp();

    uint32_t x[E];
#pragma unroll
    for (int w = 0; w < C::NW; ++w) {
        // (compile-time after unrolling)
        const int S = C::win_start(w), LB0 = C::win_lb0(w);
        const int low = tl & ((1 << S) - 1), high = tl >> S;
        const int base = low + (high << (S + LOGE));
        if (w > 0) {
            if (S <= 5) __syncwarp(); else __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < E; ++i) x[i] = sm[phys(base + (i << S))];
        switch (w) {   // window parameters must be template arguments
            case 0: fwd_window<LOGE, C::win_start(0), C::win_lb0(0), LAZY>(x, tw, low, q); break;
            case 1: fwd_window<LOGE, C::win_start(1), C::win_lb0(1), LAZY>(x, tw, low, q); break;
            case 2: fwd_window<LOGE, C::win_start(2), C::win_lb0(2), LAZY>(x, tw, low, q); break;
            default: fwd_window<LOGE, C::win_start(3), C::win_lb0(3), LAZY>(x, tw, low, q); break;
        }
        if (w + 1 < C::NW) {
#pragma unroll
            for (int i = 0; i < E; ++i) sm[phys(base + (i << S))] = x[i];
        } else if (active) {
            // last window: S = LOGN - LOGE, high = 0; lanes hold consecutive columns -> coalesced stores
#pragma unroll
            for (int i = 0; i < E; ++i) {
                uint32_t v = x[i];
                if (LAZY) v = ab_csub(v, 2 * q);
                if (LAZY) v = ab_csub(v, q);
                __stcs(yp + base + (i << S), (int64_t)v);
            }
        }
    }
}

template <int LOGN, int LOGE, bool LAZY>
__global__ void __launch_bounds__(Pow2Cfg<LOGN, LOGE>::THREADS, Pow2Cfg<LOGN, LOGE>::MINB)
k_crt_pow2_inv(const Pow2Args args, int64_t *__restrict__ y) {
    using C = Pow2Cfg<LOGN, LOGE>;
    constexpr int E = C::E, T = C::T, G = C::G;
    extern __shared__ uint32_t smem[];
    const int tid = threadIdx.x, tl = tid & (T - 1), slot = tid / T;
    int64_t trans = (int64_t)blockIdx.x * C::TPC + slot;
    const bool active = trans < args.ntrans;
    if (!active) trans = 0;
    const Pow2Limb lp = args.limb[trans % args.nlimbs];
    const uint32_t q = lp.q;
    const uint2 *__restrict__ tw = lp.tw;
    uint32_t *sm = smem + slot * C::PADN;
    int64_t *yp = y + trans * C::N;

    uint32_t x[E];
#pragma unroll
    for (int w = C::NW - 1; w >= 0; --w) {
        const int S = C::win_start(w), LB0 = C::win_lb0(w);
        const int low = tl & ((1 << S) - 1), high = tl >> S;
        const int base = low + (high << (S + LOGE));
        if (w == C::NW - 1) {
            // first window of the inverse reads HBM directly: lanes hold consecutive columns
#pragma unroll
            for (int i = 0; i < E; ++i) x[i] = active ? (uint32_t)__ldcs(yp + base + (i << S)) : 0u;
        } else {
            if (C::win_start(w + 1) <= 5) __syncwarp(); else __syncthreads();
#pragma unroll
            for (int i = 0; i < E; ++i) x[i] = sm[phys(base + (i << S))];
        }
        switch (w) {
            case 0: inv_window<LOGE, C::win_start(0), C::win_lb0(0), LAZY>(x, tw, low, q, lp.ninv, lp.ninv_s); break;
            case 1: inv_window<LOGE, C::win_start(1), C::win_lb0(1), LAZY>(x, tw, low, q, lp.ninv, lp.ninv_s); break;
            case 2: inv_window<LOGE, C::win_start(2), C::win_lb0(2), LAZY>(x, tw, low, q, lp.ninv, lp.ninv_s); break;
            default: inv_window<LOGE, C::win_start(3), C::win_lb0(3), LAZY>(x, tw, low, q, lp.ninv, lp.ninv_s); break;
        }
#pragma unroll
        for (int i = 0; i < E; ++i) sm[phys(base + (i << S))] = x[i];
    }
    __syncwarp();
    if (active) {
        const int wb = tl / G, lg = tl % G;
        int4 *dst = reinterpret_cast<int4 *>(yp + (int64_t)wb * G * E);
#pragma unroll
        for (int j = 0; j < E / 2; ++j) {
            const int a = wb * G * E + 2 * (lg + j * G);
            uint32_t v0 = sm[phys(a)], v1 = sm[phys(a + 1)];
            if (LAZY) { v0 = ab_csub(v0, q); v1 = ab_csub(v1, q); }
            __stcs(dst + lg + j * G, make_int4((int)v0, 0, (int)v1, 0));
        }
    }
}

// ---------------------------------------------------------------------------------------------
template <int LOGN, int LOGE, bool LAZY>
static void launch_one(const Pow2Args &args, int64_t *y, bool inverse, cudaStream_t st) {
    using C = Pow2Cfg<LOGN, LOGE>;
    const size_t smem = (size_t)C::TPC * C::PADN * sizeof(uint32_t);
    const unsigned grid = (unsigned)((args.ntrans + C::TPC - 1) / C::TPC);
    if (inverse) {
        auto k = k_crt_pow2_inv<LOGN, LOGE, LAZY>;
        if (smem > 48 * 1024) AB_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, C::THREADS, smem, st>>>(args, y);
    } else {
        auto k = k_crt_pow2_fwd<LOGN, LOGE, LAZY>;
        if (smem > 48 * 1024) AB_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, C::THREADS, smem, st>>>(args, y);
    }
    ab_count_launch();
    ab_check_launch("crt_pow2");
}

template <bool LAZY>
static void dispatch_logn(int logn, const Pow2Args &args, int64_t *y, bool inverse, cudaStream_t st) {
    switch (logn) {
        case 5: launch_one<5, 5, LAZY>(args, y, inverse, st); break;
        case 6: launch_one<6, 3, LAZY>(args, y, inverse, st); break;
        case 7: launch_one<7, 4, LAZY>(args, y, inverse, st); break;
        case 8: launch_one<8, 4, LAZY>(args, y, inverse, st); break;
        case 9: launch_one<9, 5, LAZY>(args, y, inverse, st); break;
        case 10: launch_one<10, 5, LAZY>(args, y, inverse, st); break;
        case 11: launch_one<11, 5, LAZY>(args, y, inverse, st); break;
        case 12: launch_one<12, 5, LAZY>(args, y, inverse, st); break;
        case 13: launch_one<13, 5, LAZY>(args, y, inverse, st); break;
        case 14: launch_one<14, 5, LAZY>(args, y, inverse, st); break;
        case 15: launch_one<15, 5, LAZY>(args, y, inverse, st); break;
        default: throw ab_error{AB_ERR_UNSUPPORTED, "pow2 kernel: unsupported size"};
    }
}

bool ab_pow2_supported(const ab_plan *p) { return p->pow2 && p->has_crt && p->logn >= 5 && p->logn <= 15; }

void ab_launch_crt_pow2(const ab_plan *p, const AbView &v, bool inverse, cudaStream_t st) {
    AB_REQUIRE(ab_pow2_supported(p), AB_ERR_UNSUPPORTED, "pow2 kernel: plan not supported");
    AB_REQUIRE(v.layout == AB_LAYOUT_PLANAR || v.nlimbs == 1, AB_ERR_UNSUPPORTED, "pow2 kernel: planar layout only");
    AB_REQUIRE(((uintptr_t)v.y & 15) == 0, AB_ERR_ARG, "buffer must be 16-byte aligned");
    if (v.batch == 0) return;
    Pow2Args args{};
    args.nlimbs = v.nlimbs;
    args.ntrans = v.batch * v.nlimbs;
    bool lazy = true;
    for (int l = 0; l < v.nlimbs; ++l) {
        const AbLimbDev &ld = p->limb[v.limb0 + l];
        args.limb[l].q = ld.q;
        args.limb[l].ninv = ld.ninv;
        args.limb[l].ninv_s = ld.ninv_s;
        args.limb[l].tw = reinterpret_cast<const uint2 *>(inverse ? ld.tw_inv : ld.tw_fwd);
        if (ld.q >= (1u << 30)) lazy = false;
    }
    if (lazy) dispatch_logn<true>(p->logn, args, v.y, inverse, st);
    else dispatch_logn<false>(p->logn, args, v.y, inverse, st);
}
