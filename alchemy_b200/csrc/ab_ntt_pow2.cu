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
// touch only the 1024-element block owned by one warp, so they need __syncwarp only.
//
// Memory pipeline (forward, n = 32768, where only one CTA fits an SM): CTAs are persistent.  As soon as every thread has pulled the
// inputs of the last window out of shared memory, the buffer is dead, and the low 32-bit halves of the
// NEXT transform's int64 residues are streamed into it with 4-byte cp.async (LDGSTS) while the last
// window's butterflies run and its results are stored: HBM reads, HBM writes and arithmetic overlap
// inside one CTA even at n = 32768 where only one CTA fits an SM.
// Round 2: from n = 8192 up the first threads of a CTA request the input of a CTA a quarter of a resident wave later into L2
// (cp.async.bulk.prefetch.L2), and the inverses at n = 8192 ... 32768 run as TWO independent CTAs per transform (k_crt_pow2_inv_h2: the
// first inverse stage is the only one that crosses the halves; each CTA computes its own output of it while loading).  The forward at
// n = 32768 is a 2-CTA cluster (its cross-half stage is the last one, so it needs the exchange over DSMEM).
// HBM traffic is one read and one write of the int64 residues: 16 n bytes per transform, the
// algorithmic minimum (the split inverses read the input twice, the second time mostly from L2: +7.5 % DRAM reads at n = 32768).  Twiddles (Shoup pairs) are window-packed so one 16-byte load fetches two;
// window 0's 31 twiddles are the same for every thread and travel as kernel constants.
#include "ab_internal.h"
#include <cooperative_groups.h>

#include <utility>

#include "ab_modarith.cuh"

#include "ab_ntt_pow2_dev.cuh"


#ifndef AB_POW2_CLUSTER
#define AB_POW2_CLUSTER 1     // n = 32768: 2-CTA cluster kernels (0: single-CTA persistent / plain kernels)
#endif
#ifndef AB_POW2_RELAXED_ARRIVE
#define AB_POW2_RELAXED_ARRIVE 1   // forward cluster kernel: the arrival that only signals "my remote reads are consumed" carries no data:
                                   // barrier.cluster.arrive.relaxed instead of the default .release, which ptxas implements with MEMBAR.ALL.GPU
                                   // (measured: forward 0.520 -> 0.497 ms per launch, 0.64 -> 0.67 of the HBM roofline)
#endif

template <int LOGN, int LOGE>
struct Pow2Cfg {
    static constexpr int N = 1 << LOGN, E = 1 << LOGE, T = N / E;
    static constexpr int TPC = (T >= 128) ? 1 : 128 / T;
    static constexpr int THREADS = T * TPC;
    static constexpr int PADN = N + (N >> 5);
    static constexpr int NW = ab_pow2_nw(LOGN, LOGE);
    static constexpr int G = (T < 32) ? T : 32;          // lanes of one warp working on one transform
    static constexpr int MINB = (1024 / THREADS) > 0 ? (1024 / THREADS) : 1;
    // persistent + cp.async prefetch pays only where a single CTA fits an SM (measured on B200: n = 32768)
    static constexpr bool PIPE = (LOGN >= 15);
};


__device__ __forceinline__ void cp_async4(uint32_t *smem_dst, const void *gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// window dispatch: the window parameters must be template arguments (w is a constant after unrolling)
#define AB_FW(W) fwd_window<LOGE, W, ab_pow2_s(LOGN, LOGE, W), ab_pow2_lb0(LOGN, LOGE, W), LAZY>(x, lp, tw, ab_pow2_base(LOGN, LOGE, W), low, q)
#define AB_IW(W) inv_window<LOGE, W, ab_pow2_s(LOGN, LOGE, W), ab_pow2_lb0(LOGN, LOGE, W), LAZY>(x, lp, tw, ab_pow2_base(LOGN, LOGE, W), low, q)
#define AB_RUN_FWD_WINDOW(w) switch (w) { case 0: AB_FW(0); break; case 1: AB_FW(1); break; case 2: AB_FW(2); break; default: AB_FW(3); break; }
#define AB_RUN_INV_WINDOW(w) switch (w) { case 0: AB_IW(0); break; case 1: AB_IW(1); break; case 2: AB_IW(2); break; default: AB_IW(3); break; }

// ---------------------------------------------------------------------------------------------
// forward, small transforms (several per CTA): load, transform, store
template <int LOGN, int LOGE, bool LAZY>
__global__ void __launch_bounds__(Pow2Cfg<LOGN, LOGE>::THREADS, Pow2Cfg<LOGN, LOGE>::MINB)
k_crt_pow2_fwd(const __grid_constant__ Pow2Args args, int64_t *__restrict__ y) {
    using C = Pow2Cfg<LOGN, LOGE>;
    constexpr int E = C::E, T = C::T, G = C::G;
    extern __shared__ uint32_t smem[];
    // grid = (groups of TPC ring elements, limbs): every slot of a CTA works in the same limb, so the limb's
    // constants are CTA-uniform (constant-bank operands, no per-thread copies)
    const int tid = threadIdx.x, tl = tid & (T - 1), slot = tid / T;
    int64_t b = (int64_t)blockIdx.x * C::TPC + slot;
    const bool active = b < args.batch;
    if (!active) b = 0;
    const Pow2Limb &lp = args.limb[blockIdx.y];
    const uint32_t q = lp.q;
    const uint2 *__restrict__ tw = lp.tw;
    uint32_t *sm = smem + slot * C::PADN;
    int64_t *yp = y + (b * args.nlimbs + blockIdx.y) * C::N;
    if constexpr (C::TPC == 1 && LOGN >= 13) pow2_prefetch<LOGN>(args, y);

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
        const int S = ab_pow2_s(LOGN, LOGE, w);
        const int low = tl & ((1 << S) - 1), high = tl >> S;
        const int base = low + (high << (S + LOGE));
        if (w > 0) {
            if (S <= 5) __syncwarp(); else __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < E; ++i) x[i] = sm[phys(base + (i << S))];
        AB_RUN_FWD_WINDOW(w)
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

// forward, large transforms: persistent CTA, next transform prefetched with cp.async under the last window
template <int LOGN, int LOGE, bool LAZY>
__global__ void __launch_bounds__(Pow2Cfg<LOGN, LOGE>::THREADS, Pow2Cfg<LOGN, LOGE>::MINB)
k_crt_pow2_fwd_pipe(const __grid_constant__ Pow2Args args, int64_t *__restrict__ y) {
    using C = Pow2Cfg<LOGN, LOGE>;
    constexpr int E = C::E, G = C::G;
    static_assert(C::TPC == 1 && G == 32, "pipelined kernel: one transform per CTA");
    extern __shared__ uint32_t sm[];
    const int tl = threadIdx.x;
    const int wb = tl / G, lg = tl % G;

    // low halves of the int64 residues of the 1024-element block this warp owns in window 0
    const Pow2Limb &lp = args.limb[blockIdx.y];
    const uint32_t q = lp.q;
    const uint2 *__restrict__ tw = lp.tw;
    auto prefetch = [&](int64_t b) {
        const int64_t *src = y + (b * args.nlimbs + blockIdx.y) * C::N + wb * G * E + lg;
        uint32_t *dst = sm + wb * G * E + wb * E + lg;      // phys(wb*1024 + lg + j*32) = that + wb*32 + j
#pragma unroll
        for (int j = 0; j < E; ++j) cp_async4(dst + j * (G + 1), src + j * G);
        cp_async_commit();
    };

    int64_t b = blockIdx.x;
    if (b < args.batch) prefetch(b);
    for (; b < args.batch; b += gridDim.x) {
        int64_t *yp = y + (b * args.nlimbs + blockIdx.y) * C::N;
        cp_async_wait_all();
        __syncwarp();
        uint32_t x[E];
#pragma unroll
        for (int w = 0; w < C::NW; ++w) {
            const int S = ab_pow2_s(LOGN, LOGE, w);
            const int low = tl & ((1 << S) - 1), high = tl >> S;
            const int base = low + (high << (S + LOGE));
            if (w > 0) {
                if (S <= 5) __syncwarp(); else __syncthreads();
            }
#pragma unroll
            for (int i = 0; i < E; ++i) x[i] = sm[phys(base + (i << S))];
            if (w + 1 == C::NW) {
                // every thread holds its last-window inputs: the buffer is dead, refill it for the next transform
                __syncthreads();
                const int64_t next = b + gridDim.x;
                if (next < args.batch) prefetch(next);
            }
            AB_RUN_FWD_WINDOW(w)
            if (w + 1 < C::NW) {
#pragma unroll
                for (int i = 0; i < E; ++i) sm[phys(base + (i << S))] = x[i];
            } else {
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
}

template <int LOGN, int LOGE, bool LAZY>
__global__ void __launch_bounds__(Pow2Cfg<LOGN, LOGE>::THREADS, Pow2Cfg<LOGN, LOGE>::MINB)
k_crt_pow2_inv(const __grid_constant__ Pow2Args args, int64_t *__restrict__ y) {
    using C = Pow2Cfg<LOGN, LOGE>;
    constexpr int E = C::E, T = C::T, G = C::G;
    extern __shared__ uint32_t smem[];
    const int tid = threadIdx.x, tl = tid & (T - 1), slot = tid / T;
    int64_t b = (int64_t)blockIdx.x * C::TPC + slot;
    const bool active = b < args.batch;
    if (!active) b = 0;
    const Pow2Limb &lp = args.limb[blockIdx.y];
    const uint32_t q = lp.q;
    const uint2 *__restrict__ tw = lp.tw;
    uint32_t *sm = smem + slot * C::PADN;
    int64_t *yp = y + (b * args.nlimbs + blockIdx.y) * C::N;
    if constexpr (C::TPC == 1 && LOGN >= 13) pow2_prefetch<LOGN>(args, y);

    uint32_t x[E];
#pragma unroll
    for (int w = C::NW - 1; w >= 0; --w) {
        const int S = ab_pow2_s(LOGN, LOGE, w);
        const int low = tl & ((1 << S) - 1), high = tl >> S;
        const int base = low + (high << (S + LOGE));
        if (w == C::NW - 1) {
            // first window of the inverse reads HBM directly: lanes hold consecutive columns
#pragma unroll
            for (int i = 0; i < E; ++i) x[i] = active ? (uint32_t)__ldcs(yp + base + (i << S)) : 0u;
        } else {
            if (ab_pow2_s(LOGN, LOGE, w + 1) <= 5) __syncwarp(); else __syncthreads();
#pragma unroll
            for (int i = 0; i < E; ++i) x[i] = sm[phys(base + (i << S))];
        }
        AB_RUN_INV_WINDOW(w)
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
// n = 32768 as a 2-CTA cluster: each CTA (512 threads) owns half of the transform (16 blocks of 1024 residues,
// 67.6 KB of shared memory), so TWO half-transforms of different clusters are resident per SM and their load,
// butterfly and store phases overlap -- what a single 1024-thread CTA filling the SM cannot do.  Windows 0 and 1
// stay inside 1024-residue blocks (warp-local); only the last window (stride 1024 ... 16384) crosses the halves:
// its 32 inputs per thread come half from local and half from the peer CTA's shared memory (DSMEM).
namespace cg = cooperative_groups;

template <int LOGE, bool LAZY>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(512, 2)
k_crt_pow2_fwd_c2(const __grid_constant__ Pow2Args args, int64_t *__restrict__ y) {
    constexpr int LOGN = 15, E = 1 << LOGE, G = 32;
    static_assert(LOGE == 5, "cluster kernel: 3 windows of 5 stages");
    extern __shared__ uint32_t sm[];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned r = cluster.block_rank();
    const int tl = threadIdx.x, wb = tl / G, lg = tl % G;
    const int64_t b = blockIdx.x >> 1;
    const Pow2Limb &lp = args.limb[blockIdx.y];
    const uint32_t q = lp.q;
    const uint2 *__restrict__ tw = lp.tw;
    int64_t *yp = y + (b * args.nlimbs + blockIdx.y) * (1 << LOGN);
    pow2_c2_prefetch<15>(args, y);
    {   // this warp's 1024-residue block of this CTA's half
        const int4 *src = reinterpret_cast<const int4 *>(yp + (int64_t)(16 * r + wb) * G * E);
#pragma unroll
        for (int j0 = 0; j0 < E / 2; j0 += 8) {
            int4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __ldcs(src + lg + (j0 + j) * G);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int a = wb * G * E + 2 * (lg + (j0 + j) * G);
                sm[phys(a)] = (uint32_t)v[j].x;
                sm[phys(a + 1)] = (uint32_t)v[j].z;
            }
        }
    }
    __syncwarp();
    uint32_t x[E];
#pragma unroll
    for (int w = 0; w < 2; ++w) {     // stages 0..9: inside the blocks of this half (local indices)
        const int S = ab_pow2_s(LOGN, LOGE, w);
        const int low = tl & ((1 << S) - 1), high = tl >> S;
        const int base = low + (high << (S + LOGE));
        if (w > 0) __syncwarp();
#pragma unroll
        for (int i = 0; i < E; ++i) x[i] = sm[phys(base + (i << S))];
        AB_RUN_FWD_WINDOW(w)
#pragma unroll
        for (int i = 0; i < E; ++i) sm[phys(base + (i << S))] = x[i];
    }
    cluster.sync();      // both halves are complete (and both CTAs are running: first access to the peer's shared memory)
    {
        const int low = 512 * (int)r + tl;            // column of the last window, 0..1023
        const uint32_t *peer = cluster.map_shared_rank(sm, r ^ 1);
        const uint32_t *h0 = (r == 0) ? sm : peer, *h1 = (r == 0) ? peer : sm;
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = h0[phys(i * 1024 + low)];
#pragma unroll
        for (int i = 0; i < 16; ++i) x[16 + i] = h1[phys(i * 1024 + low)];
#if AB_POW2_RELAXED_ARRIVE
        AB_FW(2);
        // every remote value has been consumed by the butterflies above (data dependence), so a relaxed arrival suffices
        asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
#else
        cluster.barrier_arrive();                     // my remote reads are done (waited for before exit)
        AB_FW(2);
#endif
#pragma unroll
        for (int i = 0; i < E; ++i) {
            uint32_t v = x[i];
            if (LAZY) v = ab_csub(v, 2 * q);
            if (LAZY) v = ab_csub(v, q);
            __stcs(yp + low + (i << 10), (int64_t)v);
        }
        cluster.barrier_wait();                       // the peer no longer reads my shared memory
    }
}

// n = 32768 (and n = 16384, 8192) inverse as two independent CTAs per transform, WITHOUT a cluster (round 2; the 2-CTA cluster version it replaces pushed half of the first window's outputs into the
// peer's shared memory and measured 0.566 ms per 2.147 GB launch against 0.515 ms here): the first stage of the inverse (stride 16384, one twiddle per butterfly) is the only one that
// crosses the halves, and after it the two halves are independent 16384-point transforms.  So CTA r of a pair computes just ITS output
// of every stride-16384 butterfly while loading -- r = 0 the sums, r = 1 the twiddled differences -- then stages 13..10 on its 16 rows of
// the 32 x 1024 view, two columns per thread, and the remaining two windows in its own shared memory.  Both CTAs read the whole input
// (the second reader hits L2; HBM traffic stays the algorithmic 16 bytes per residue), nothing crosses SMs, no cluster barrier and no
// MEMBAR; the multiplications are the same in total (CTA 1 does the 16384 of stage 14, CTA 0 none).
template <int LOGN, int LOGE, bool LAZY>
__global__ void __launch_bounds__(1 << (LOGN - 6), 1024 >> (LOGN - 6))
k_crt_pow2_inv_h2(const __grid_constant__ Pow2Args args, int64_t *__restrict__ y) {
    // the transform as 32 rows x COLS columns; this CTA: rows 16 r .. 16 r + 15, every column, two columns per thread
    constexpr int E = 1 << LOGE, G = 32, S2 = LOGN - 5, COLS = 1 << S2, T = COLS / 2, LB0 = ab_pow2_lb0(LOGN, LOGE, 2), KOFF = LB0 ? (1 << LB0) : 0;
    static_assert(LOGE == 5 && LOGN >= 13 && LOGN <= 15, "half kernel: top stage, a window of up to 4 stages, 2 windows of 5 stages");
    extern __shared__ uint32_t sm[];
    const unsigned r = blockIdx.x & 1;
    const int tl = threadIdx.x, wb = tl / G, lg = tl % G;
    const int64_t b = blockIdx.x >> 1;
    const Pow2Limb &lp = args.limb[blockIdx.y];
    const uint32_t q = lp.q;
    const uint2 *__restrict__ tw = lp.tw;
    int64_t *yp = y + (b * args.nlimbs + blockIdx.y) * (1 << LOGN);
    pow2_c2_prefetch<LOGN>(args, y);
    {
        constexpr int64_t base2 = ab_pow2_base(LOGN, LOGE, 2);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int col = tl + T * c;
            uint32_t lo[16], hi[16], x16[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                lo[i] = (uint32_t)__ldcs(yp + col + (i << S2));
                hi[i] = (uint32_t)__ldcs(yp + col + ((i + 16) << S2));
            }
            if (r == 0) {
#pragma unroll
                for (int i = 0; i < 16; ++i) x16[i] = LAZY ? ab_csub(lo[i] + hi[i], 2 * q) : ab_addmod(lo[i], hi[i], q);
            } else {
                // twiddle k = 16 + i (- KOFF) of window 2 for this column: entries (k >> 1) << S2 of the packed table, two per 16-byte load
                const uint4 *tp = reinterpret_cast<const uint4 *>(tw + base2) + col;
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const uint4 pr = __ldg(tp + ((int64_t)((16 + i - KOFF) >> 1) << S2));
                    x16[i] = LAZY ? ab_shoup_lazy(lo[i] - hi[i] + 2 * q, pr.x, pr.y, q) : ab_shoup(ab_submod(lo[i], hi[i], q), pr.x, pr.y, q);
                    x16[i + 1] = LAZY ? ab_shoup_lazy(lo[i + 1] - hi[i + 1] + 2 * q, pr.z, pr.w, q) : ab_shoup(ab_submod(lo[i + 1], hi[i + 1], q), pr.z, pr.w, q);
                }
            }
            inv_window<4, 2, S2, LB0, LAZY>(x16, lp, tw, base2, col, q);      // the other stages of window 2: its first 16 twiddles
#pragma unroll
            for (int i = 0; i < 16; ++i) sm[phys(i * COLS + col)] = x16[i];
        }
    }
    __syncthreads();
    uint32_t x[E];
#pragma unroll
    for (int w = 1; w >= 0; --w) {
        const int S = ab_pow2_s(LOGN, LOGE, w);
        const int low = tl & ((1 << S) - 1), high = tl >> S;
        const int base = low + (high << (S + LOGE));
        if (w < 1) __syncwarp();
#pragma unroll
        for (int i = 0; i < E; ++i) x[i] = sm[phys(base + (i << S))];
        AB_RUN_INV_WINDOW(w)
#pragma unroll
        for (int i = 0; i < E; ++i) sm[phys(base + (i << S))] = x[i];
    }
    __syncwarp();
    {
        int4 *dst = reinterpret_cast<int4 *>(yp + (int64_t)((T / G) * r + wb) * G * E);
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
template <class K>
static int resident_ctas(K kernel, int threads, size_t smem) {
    int nb = 0;
    AB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, smem));
    return nb < 1 ? 1 : nb;
}

static bool pow2_split() {      // n = 8192 and 16384 inverse as two independent half-transform CTAs too: 0.775 -> 0.810 and 0.693 -> 0.720 of the HBM
    static const bool on = !(getenv("AB_POW2_SPLIT") && atoi(getenv("AB_POW2_SPLIT")) == 0);      // roofline (AB_POW2_SPLIT=0: one CTA per transform)
    return on;
}

template <int LOGN, int LOGE, bool LAZY>
static void launch_one(const ab_plan *p, const Pow2Args &args, int64_t *y, bool inverse, cudaStream_t st) {
    using C = Pow2Cfg<LOGN, LOGE>;
    const size_t smem = (size_t)C::TPC * C::PADN * sizeof(uint32_t);
    const dim3 grid((unsigned)((args.batch + C::TPC - 1) / C::TPC), (unsigned)args.nlimbs);
    // L2 prefetch distance of the large transforms, in resident waves of CTAs (0.125 ... 0.5 measured the same at n = 32768: 0.612 -> 0.629 of
    // the HBM roofline; one wave ahead gains half of that, and beyond the lines fall out of L2 again); 0 switches it off
    static const double pf_waves = getenv("AB_POW2_PF_WAVES") ? atof(getenv("AB_POW2_PF_WAVES")) : 0.25;
    Pow2Args pa = args;
    pa.pf_dist = (int32_t)(pf_waves * p->sm_count * ((LOGN == 15 && AB_POW2_CLUSTER) ? 2 : C::MINB));
    if constexpr (LOGN == 15 && AB_POW2_CLUSTER) {
        constexpr size_t half = ((size_t)(1 << 14) + (1 << 9)) * sizeof(uint32_t);
        static thread_local bool ready = false;
        if (!ready) {
            AB_CHECK_CUDA(cudaFuncSetAttribute(k_crt_pow2_fwd_c2<LOGE, LAZY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)half));
            AB_CHECK_CUDA(cudaFuncSetAttribute(k_crt_pow2_inv_h2<LOGN, LOGE, LAZY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)half));
            ready = true;
        }
        const dim3 g2((unsigned)(2 * args.batch), (unsigned)args.nlimbs);
        if (inverse) k_crt_pow2_inv_h2<LOGN, LOGE, LAZY><<<g2, 512, half, st>>>(pa, y);
        else k_crt_pow2_fwd_c2<LOGE, LAZY><<<g2, 512, half, st>>>(pa, y);
    } else if (inverse && (LOGN == 14 || LOGN == 13) && pow2_split()) {
        if constexpr (LOGN == 14 || LOGN == 13) {
            constexpr size_t half = ((size_t)(1 << (LOGN - 1)) + (1 << (LOGN - 6))) * sizeof(uint32_t);
            static thread_local bool ready = false;
            if (!ready) {
                AB_CHECK_CUDA(cudaFuncSetAttribute(k_crt_pow2_inv_h2<LOGN, LOGE, LAZY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)half));
                ready = true;
            }
            pa.pf_dist = (int32_t)(pf_waves * p->sm_count * (1024 >> (LOGN - 6)));
            k_crt_pow2_inv_h2<LOGN, LOGE, LAZY><<<dim3((unsigned)(2 * args.batch), (unsigned)args.nlimbs), 1 << (LOGN - 6), half, st>>>(pa, y);
        }
    } else if (inverse) {
        auto k = k_crt_pow2_inv<LOGN, LOGE, LAZY>;
        static thread_local bool ready = false;
        if (!ready && smem > 48 * 1024) AB_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ready = true;
        k<<<grid, C::THREADS, smem, st>>>(pa, y);
    } else if constexpr (C::PIPE) {
        auto k = k_crt_pow2_fwd_pipe<LOGN, LOGE, LAZY>;
        static thread_local int per_sm = 0;
        if (!per_sm) {
            if (smem > 48 * 1024) AB_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            per_sm = resident_ctas(k, C::THREADS, smem);
        }
        // persistent: about one resident wave, split evenly over the limbs
        const int64_t gx = std::max<int64_t>(1, std::min<int64_t>(args.batch, ((int64_t)p->sm_count * per_sm) / args.nlimbs));
        k<<<dim3((unsigned)gx, (unsigned)args.nlimbs), C::THREADS, smem, st>>>(args, y);
    } else {
        auto k = k_crt_pow2_fwd<LOGN, LOGE, LAZY>;
        static thread_local bool ready = false;
        if (!ready && smem > 48 * 1024) AB_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ready = true;
        k<<<grid, C::THREADS, smem, st>>>(pa, y);
    }
    ab_count_launch();
    ab_check_launch("crt_pow2");
}

template <bool LAZY>
static void dispatch_logn(const ab_plan *p, const Pow2Args &args, int64_t *y, bool inverse, cudaStream_t st) {
#define AB_CASE(LN) case LN: launch_one<LN, ab_pow2_loge(LN), LAZY>(p, args, y, inverse, st); break;
    switch (p->logn) {
        AB_CASE(5) AB_CASE(6) AB_CASE(7) AB_CASE(8) AB_CASE(9) AB_CASE(10) AB_CASE(11) AB_CASE(12) AB_CASE(13) AB_CASE(14) AB_CASE(15)
        default: throw ab_error{AB_ERR_UNSUPPORTED, "pow2 kernel: unsupported size"};
    }
#undef AB_CASE
}

bool ab_pow2_supported(const ab_plan *p) { return p->pow2 && p->has_crt && p->logn >= 5 && p->logn <= 15; }

void ab_launch_crt_pow2(const ab_plan *p, const AbView &v, bool inverse, cudaStream_t st) {
    AB_REQUIRE(ab_pow2_supported(p), AB_ERR_UNSUPPORTED, "pow2 kernel: plan not supported");
    AB_REQUIRE(v.layout == AB_LAYOUT_PLANAR || v.nlimbs == 1, AB_ERR_UNSUPPORTED, "pow2 kernel: planar layout only");
    AB_REQUIRE(((uintptr_t)v.y & 15) == 0, AB_ERR_ARG, "buffer must be 16-byte aligned");
    if (v.batch == 0) return;
    Pow2Args args{};
    args.nlimbs = v.nlimbs;
    args.batch = v.batch;
    bool lazy = true;
    for (int l = 0; l < v.nlimbs; ++l) {
        const AbLimbDev &ld = p->limb[v.limb0 + l];
        args.limb[l].q = ld.q;
        args.limb[l].ninv = ld.ninv;
        args.limb[l].ninv_s = ld.ninv_s;
        args.limb[l].tw = reinterpret_cast<const uint2 *>(inverse ? ld.tw_inv : ld.tw_fwd);
        const uint2_ *t0 = inverse ? ld.tw0_inv : ld.tw0_fwd;
        for (int k = 0; k < 32; ++k) args.limb[l].tw0[k] = make_uint2(t0[k].x, t0[k].y);
        if (ld.q >= (1u << 30)) lazy = false;
    }
    if (lazy) dispatch_logn<true>(p, args, v.y, inverse, st);
    else dispatch_logn<false>(p, args, v.y, inverse, st);
}
