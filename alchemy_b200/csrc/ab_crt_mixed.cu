// ab_crt_mixed.cu -- batched CRT / crtInv for arbitrary cyclotomic index m, prime moduli below 2^31 (sm_100a).
//
// Replaces lol-cpp `tensorCRTRq` / `tensorCRTInvRq` [RECALL] on the non-power-of-two rings of the reference's
// examples (examples/Common.hs:41-54: m' = 11648 ... 54600) and of BASELINE config 3 (m = 40320).
// One CTA owns one (ring element, limb): the n = phi(m) int64 residues are read once from HBM with 16-byte
// loads and narrowed to uint32 in shared memory, the program of ab_mixed_plan.cpp runs out of shared memory,
// and the result is written once: 16 n bytes of HBM traffic per transform, the algorithmic minimum.
//
// A dense step computes out_r = sum_c M[r][c] x_c for P <= 12 strided residues held in registers.  The
// constants are stored as M * 2^32 mod q, the sum is accumulated exactly in 64 bits (IMAD.WIDE, one
// instruction per term) and reduced ONCE per output with a Montgomery step: P + 4 instructions per output
// instead of 7 per term.  The limb's constant tables (a few KB) sit in shared memory and are fetched with
// 16-byte broadcast loads.  Radix-2 stages of the 2^e axis run three at a time on 8 residues per thread.
#include <algorithm>
#include <cstdlib>
#include <vector>
#include <utility>

#include "ab_internal.h"
#include "ab_modarith.cuh"

struct MxArgs {
    AbMxStep steps[AB_MX_MAX_STEPS];
    struct {
        uint32_t q, qinv_neg;      // -q^-1 mod 2^32
        const uint32_t *tab;
        int32_t wide, pad;         // table layout of this limb (ab_mixed_plan.cpp)
    } limb[AB_MAX_LIMBS];
    int32_t nsteps, n, nlimbs;
    int32_t tab_words[2];          // narrow, wide
    int64_t batch;
    int64_t estride, lstride, bstride;   // int64 strides of coefficient, limb, ring element
    AbMxFused fx;                        // fused load / store (ab_internal.h); all zero = plain in-place transform
};

#ifndef AB_MX_MINB
#define AB_MX_MINB 4          // resident CTAs per SM the register budget is tuned for
#endif
#ifndef AB_MX_SPLIT
#define AB_MX_SPLIT 0         // 1: two partial accumulators per output (shorter dependent chain); measured: no gain on B200
#endif

template <class F, int... I>
__device__ __forceinline__ void mx_static_for_impl(F &&f, std::integer_sequence<int, I...>) {
    (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void mx_static_for(F &&f) {
    mx_static_for_impl(f, std::make_integer_sequence<int, N>{});
}

// work item -> (tb, off): V adjacent columns per item, div_magic divides by estride / V
template <int V>
__device__ __forceinline__ void item_coords(const AbMxStep &s, int v, int &tb, int &off) {
    const int es = s.estride / V;
    tb = s.div_magic ? (int)__umulhi((uint32_t)v, s.div_magic) : v;
    off = (v - tb * es) * V;
}
__device__ __forceinline__ int item_row(const AbMxStep &s, int tb, int off) {
    if (s.row_mode == 1) return tb - (int)__umulhi((uint32_t)tb, s.row_magic) * s.rows;
    if (s.row_mode == 2) return s.row_magic ? (int)__umulhi((uint32_t)off, s.row_magic) : off;
    return 0;
}
template <int V>
__device__ __forceinline__ void ld_cols(const uint32_t *p, uint32_t (&x)[V]) {
    if (V == 2) { const uint2 t = *reinterpret_cast<const uint2 *>(p); x[0] = t.x; x[V - 1] = t.y; }
    else x[0] = *p;
}
template <int V>
__device__ __forceinline__ void st_cols(uint32_t *p, const uint32_t (&x)[V]) {
    if (V == 2) *reinterpret_cast<uint2 *>(p) = make_uint2(x[0], x[V - 1]);
    else *p = x[0];
}

// out = M x for V columns with one Montgomery reduction per output; x in [0,q), result in [0,q).
// narrow (q < 2^30): T = sum x_c M'_c < P q^2, T + m q < 2^64, t < 4q.
// WIDE  (q < 2^31): x = xh 2^16 + xl, T = sum xl_c M'_c + xh_c M''_c with M'' = M' 2^16: every term < 2^47.
template <int P, int V, bool WIDE>
__device__ __forceinline__ void matvec_item(uint32_t *sm, const uint32_t *__restrict__ tab, const AbMxStep &s, int v, uint32_t q, uint32_t qinv_neg) {
    constexpr int PP = (P + 3) & ~3, RW = WIDE ? 2 * PP : PP;
    int tb, off;
    item_coords<V>(s, v, tb, off);
    const int es = s.estride;
    uint32_t *e = sm + tb * (P * es) + off;
    const uint4 *M = reinterpret_cast<const uint4 *>(tab + (WIDE ? s.tab_off_w : s.tab_off) + item_row(s, tb, off) * (P * RW));
    uint32_t x[P][V];
    mx_static_for<P>([&](auto C) { ld_cols<V>(e + decltype(C)::value * es, x[decltype(C)::value]); });
    uint32_t xh[WIDE ? P : 1][V];
    if (WIDE) {
        mx_static_for<P>([&](auto C) {
            constexpr int c = decltype(C)::value;
#pragma unroll
            for (int k = 0; k < V; ++k) { xh[WIDE ? c : 0][k] = x[c][k] >> 16; x[c][k] &= 0xffffu; }
        });
    }
    mx_static_for<P>([&](auto R) {
        constexpr int r = decltype(R)::value;
        uint64_t acc[V], acc2[V];
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] = acc2[k] = 0;
        mx_static_for<PP / 4>([&](auto K) {
            constexpr int k4 = decltype(K)::value;
            const uint4 m4 = M[r * (RW / 4) + k4];
            const uint32_t mm[4] = {m4.x, m4.y, m4.z, m4.w};
            mx_static_for<4>([&](auto J) {
                constexpr int c = 4 * k4 + decltype(J)::value;
                if (c < P) {
#pragma unroll
                    for (int k = 0; k < V; ++k) {
                        if (AB_MX_SPLIT && (c & 1)) acc2[k] += (uint64_t)x[c % P][k] * mm[decltype(J)::value];
                        else acc[k] += (uint64_t)x[c % P][k] * mm[decltype(J)::value];
                    }
                }
            });
            if (WIDE) {
                const uint4 h4 = M[r * (RW / 4) + PP / 4 + k4];
                const uint32_t hh[4] = {h4.x, h4.y, h4.z, h4.w};
                mx_static_for<4>([&](auto J) {
                    constexpr int c = 4 * k4 + decltype(J)::value;
                    if (c < P) {
#pragma unroll
                        for (int k = 0; k < V; ++k) {
                            if (AB_MX_SPLIT && (c & 1)) acc2[k] += (uint64_t)xh[WIDE ? c % P : 0][k] * hh[decltype(J)::value];
                            else acc[k] += (uint64_t)xh[WIDE ? c % P : 0][k] * hh[decltype(J)::value];
                        }
                    }
                });
            }
        });
        uint32_t o[V];
#pragma unroll
        for (int k = 0; k < V; ++k) {
            if (AB_MX_SPLIT) acc[k] += acc2[k];
            const uint32_t mq = (uint32_t)acc[k] * qinv_neg;
            uint32_t t = (uint32_t)((acc[k] + (uint64_t)mq * q) >> 32);
            if (!WIDE && P > 4) t = ab_csub(t, 2 * q);
            o[k] = ab_csub(t, q);
        }
        st_cols<V>(e + r * es, o);      // all inputs are already in registers: in place is safe
    });
}

// G radix-2 stages on 2^G residues (V columns); forward: Cooley-Tukey (twiddle the upper input), inverse:
// Gentleman-Sande.  LAZY (q < 2^30): Harvey butterflies, values in [0,4q) / [0,2q); otherwise exact.
template <int G, int V, bool INV, bool LAZY>
__device__ __forceinline__ void radix2_item(uint32_t *sm, const uint32_t *__restrict__ tab, const AbMxStep &s, int v, uint32_t q) {
    constexpr int E = 1 << G;
    int tb, off;
    item_coords<V>(s, v, tb, off);
    const int es = s.estride;
    uint32_t *e = sm + tb * (E * es) + off;
    const int l = item_row(s, tb, off);
    const int h0 = s.rows;
    const uint2 *T = reinterpret_cast<const uint2 *>(tab + (LAZY ? s.tab_off : s.tab_off_w));
    uint32_t x[E][V];
    mx_static_for<E>([&](auto C) { ld_cols<V>(e + decltype(C)::value * es, x[decltype(C)::value]); });
    const bool last = INV && (s.flags & 1) != 0;
    const uint2 sc = T[h0 * (E - 1)];
    mx_static_for<G>([&](auto TT) {
        constexpr int t = INV ? G - 1 - decltype(TT)::value : decltype(TT)::value;
        const uint2 *Tt = T + h0 * ((1 << t) - 1) + l;
        mx_static_for<(1 << t)>([&](auto LO) {
            constexpr int lo = decltype(LO)::value;
            const uint2 w = Tt[lo * h0];
            mx_static_for<(E >> (t + 1))>([&](auto HI) {
                constexpr int c = lo + (decltype(HI)::value << (t + 1)), d = c | (1 << t);
#pragma unroll
                for (int k = 0; k < V; ++k) {
                    if (!INV) {
                        if (LAZY) {
                            const uint32_t a = ab_csub(x[c][k], 2 * q);
                            const uint32_t b = ab_shoup_lazy(x[d][k], w.x, w.y, q);
                            x[c][k] = a + b;
                            x[d][k] = a - b + 2 * q;
                        } else {
                            const uint32_t a = x[c][k], b = ab_shoup(x[d][k], w.x, w.y, q);
                            x[c][k] = ab_addmod(a, b, q);
                            x[d][k] = ab_submod(a, b, q);
                        }
                    } else {
                        const uint32_t a = x[c][k], b = x[d][k];
                        if (LAZY) {
                            uint32_t sum = ab_csub(a + b, 2 * q);
                            if (t == 0 && last) sum = ab_shoup_lazy(sum, sc.x, sc.y, q);
                            x[c][k] = sum;
                            x[d][k] = ab_shoup_lazy(a - b + 2 * q, w.x, w.y, q);
                        } else {
                            uint32_t sum = ab_addmod(a, b, q);
                            if (t == 0 && last) sum = ab_shoup(sum, sc.x, sc.y, q);
                            x[c][k] = sum;
                            x[d][k] = ab_shoup(ab_submod(a, b, q), w.x, w.y, q);
                        }
                    }
                }
            });
        });
    });
    mx_static_for<E>([&](auto C) {
        constexpr int c = decltype(C)::value;
        uint32_t o[V];
#pragma unroll
        for (int k = 0; k < V; ++k) {
            uint32_t t = x[c][k];
            if (LAZY && !INV) t = ab_csub(t, 2 * q);
            if (LAZY) t = ab_csub(t, q);
            o[k] = t;
        }
        st_cols<V>(e + c * es, o);
    });
}

template <bool INV, bool WIDE>
__device__ __forceinline__ void run_program(const MxArgs &args, uint32_t *sm, const uint32_t *tab, uint32_t q, uint32_t qinv_neg) {
    for (int si = 0; si < args.nsteps; ++si) {
        const AbMxStep &s = args.steps[si];
        const int nw = s.nwork / s.vec;
        if (s.kind == 0) {
#define AB_MV(P)                                                                                                    \
    case P:                                                                                                         \
        if (s.vec == 2) for (int v = threadIdx.x; v < nw; v += blockDim.x) matvec_item<P, 2, WIDE>(sm, tab, s, v, q, qinv_neg); \
        else for (int v = threadIdx.x; v < nw; v += blockDim.x) matvec_item<P, 1, WIDE>(sm, tab, s, v, q, qinv_neg);            \
        break;
            switch (s.span) {
                AB_MV(2) AB_MV(3) AB_MV(4) AB_MV(5) AB_MV(6) AB_MV(7) AB_MV(10) AB_MV(11) AB_MV(12)
                default: break;
            }
#undef AB_MV
        } else {
#define AB_R2(G)                                                                                                    \
    case G:                                                                                                         \
        if (s.vec == 2) for (int v = threadIdx.x; v < nw; v += blockDim.x) radix2_item<G, 2, INV, !WIDE>(sm, tab, s, v, q);     \
        else for (int v = threadIdx.x; v < nw; v += blockDim.x) radix2_item<G, 1, INV, !WIDE>(sm, tab, s, v, q);                \
        break;
            switch (s.g) {
                AB_R2(1) AB_R2(2)
                default:
                AB_R2(3)
            }
#undef AB_R2
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// gadget digit d of the powerful-basis element `su` (optionally gathered through row `krow` of an index table:
// coeffs . embed of the tunnel), centred and reduced into the modulus q of this CTA -> sm[0..n)
__device__ __forceinline__ bool mx_digit_is_trivial(const AbMxFused &fx, int d, int limb) {
    return fx.gbase == 0 && fx.triv_src != nullptr && fx.dig[d].limb == limb;
}
__device__ __forceinline__ uint32_t mx_reduce_signed(int32_t c, uint32_t q) {
    uint32_t a = c < 0 ? (uint32_t)(-c) : (uint32_t)c;
    if (a >= q) { a -= q; if (a >= q) a %= q; }     // |c| < 2q almost always (moduli of similar size)
    return (c < 0 && a) ? q - a : a;
}
__device__ __forceinline__ void mx_load_digit(const AbMxFused &fx, int d, int64_t su, int krow, int n, uint32_t q, uint32_t *sm) {
    const uint32_t qs = fx.dig[d].q_src;
    const int lastd = fx.dig[d].last, gb = fx.gbase, hb = fx.gbase / 2;
    const uint32_t pw = fx.dig[d].pw;
    const int64_t H = fx.dig[d].H;
    const int pw_shift = (gb > 0 && (gb & (gb - 1)) == 0) ? 31 - __clz((int)pw) : -1;
    const int32_t *__restrict__ gt = fx.gather ? fx.gather + (int64_t)krow * n : nullptr;
    const int64_t *__restrict__ src = fx.src + su * fx.src_ustride + (int64_t)fx.dig[d].limb * fx.n_src;
    const int T = blockDim.x;
    for (int j0 = threadIdx.x; j0 < n; j0 += 4 * T) {     // four loads in flight per thread
        int32_t g[4];
        uint32_t v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u * T;
            g[u] = j < n ? (gt ? __ldg(gt + j) : j) : -1;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = g[u] >= 0 ? (uint32_t)__ldg(src + g[u]) : 0u;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u * T;
            if (j >= n) continue;
            int32_t c = ab_center(v[u], qs);
            if (gb != 0) {
                // balanced base-b digit without iterating: ordinary digit of c + H, shifted (ab_internal.h)
                const int64_t t = (int64_t)c + H;
                int64_t qd;
                if (pw_shift >= 0) qd = t >> pw_shift;                   // power-of-two base: arithmetic shift = floor
                else if (t >= 0) qd = (int64_t)((uint32_t)t / pw);       // t < 2^32
                else { qd = t / (int64_t)pw; if (qd * (int64_t)pw > t) qd -= 1; }
                if (lastd) c = (int32_t)qd;
                else if (pw_shift >= 0) c = (int32_t)(qd & (gb - 1)) - hb;
                else c = (int32_t)(qd >= 0 ? (int64_t)((uint32_t)qd % (uint32_t)gb) : ((qd % gb) + gb) % gb) - hb;
            }
            sm[j] = mx_reduce_signed(c, q);
        }
    }
}

// Two-phase digit load for the fused inner product with a BaseBGad gadget, where the nd digits of one source limb
// all come from the same centred coefficients: phase A gathers them ONCE per (k, source limb) -- the only place
// with dependent loads (index table, then coefficient), issued four at a time -- and parks t = centred + H as
// int64 in a scratch slice owned by this CTA; phase B cuts digit d out of t with shifts (power-of-two base) or
// divisions and reduces it into this CTA's modulus.  Thread t touches the same positions in both phases.
__device__ __forceinline__ void mx_gather_centered(const AbMxFused &fx, int d0, int64_t su, int krow, int n, int64_t *__restrict__ tbuf) {
    const uint32_t qs = fx.dig[d0].q_src;
    const int64_t H = fx.dig[d0].H;
    const int32_t *__restrict__ gt = fx.gather ? fx.gather + (int64_t)krow * n : nullptr;
    const int64_t *__restrict__ src = fx.src + su * fx.src_ustride + (int64_t)fx.dig[d0].limb * fx.n_src;
    const int T = blockDim.x;
    for (int j0 = threadIdx.x; j0 < n; j0 += 4 * T) {
        int32_t g[4];
        uint32_t v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u * T;
            g[u] = j < n ? (gt ? __ldg(gt + j) : j) : -1;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = g[u] >= 0 ? (uint32_t)__ldg(src + g[u]) : 0u;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u * T;
            if (j < n) __stcg(tbuf + j, (int64_t)ab_center(v[u], qs) + H);
        }
    }
}
__device__ __forceinline__ void mx_digit_from_t(const AbMxFused &fx, int d, int n, uint32_t q, const int64_t *__restrict__ tbuf, uint32_t *sm) {
    const int lastd = fx.dig[d].last, gb = fx.gbase, hb = fx.gbase / 2;
    const uint32_t pw = fx.dig[d].pw;
    const int pw_shift = (gb > 0 && (gb & (gb - 1)) == 0) ? 31 - __clz((int)pw) : -1;
    const int T = blockDim.x;
    if (pw_shift >= 0 && !lastd && pw_shift + (31 - __clz(gb)) <= 32 && (uint32_t)hb < q) {
        // power-of-two base, inner digit: the wanted bits of floor(t / 2^s) sit in the low word of t (two's complement),
        // and |digit| <= b/2 < q needs no reduction
        const uint32_t mask = (uint32_t)gb - 1u;
        auto dig = [&](int lo) {
            const int32_t c = (int32_t)(((uint32_t)lo >> pw_shift) & mask) - hb;
            return c < 0 ? q + (uint32_t)c : (uint32_t)c;
        };
        // four consecutive positions per thread: two 16-byte loads of t, one 16-byte shared store
        const int n4 = n >> 2;
        const int4 *t4 = reinterpret_cast<const int4 *>(tbuf);
        uint4 *s4 = reinterpret_cast<uint4 *>(sm);
        for (int g = threadIdx.x; g < n4; g += T) {
            const int4 a = __ldcg(t4 + 2 * g), b = __ldcg(t4 + 2 * g + 1);
            s4[g] = make_uint4(dig(a.x), dig(a.z), dig(b.x), dig(b.z));
        }
        for (int j = 4 * n4 + threadIdx.x; j < n; j += T) sm[j] = dig((int)__ldcg(tbuf + j));
        return;
    }
    for (int j0 = threadIdx.x; j0 < n; j0 += 4 * T) {
        int64_t t[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) t[u] = (j0 + u * T < n) ? __ldcg(tbuf + j0 + u * T) : 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u * T;
            if (j < n) {
                int64_t qd;
                if (pw_shift >= 0) qd = t[u] >> pw_shift;                      // power-of-two base: arithmetic shift = floor
                else if (t[u] >= 0) qd = (int64_t)((uint32_t)t[u] / pw);       // t < 2^32
                else { qd = t[u] / (int64_t)pw; if (qd * (int64_t)pw > t[u]) qd -= 1; }
                int32_t c;
                if (lastd) c = (int32_t)qd;
                else if (pw_shift >= 0) c = (int32_t)(qd & (gb - 1)) - hb;
                else c = (int32_t)(qd >= 0 ? (int64_t)((uint32_t)qd % (uint32_t)gb) : ((qd % gb) + gb) % gb) - hb;
                sm[j] = mx_reduce_signed(c, q);
            }
        }
    }
}

// L (inverse = false: running sums) or L^-1 (differences) along the p-1 axis of every odd prime, on the residues in
// shared memory; one vector of d = p-1 strided residues per work item, a barrier after every axis
__device__ __forceinline__ void mx_l_pass(uint32_t *sm, int n, const AbMxFused &fx, uint32_t q, bool inverse) {
    for (int a = 0; a < fx.n_odd; ++a) {
        const int d = fx.odd[a].d, rts = fx.odd[a].rts, nvec = n / d;
        for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
            const int t = v / rts, r = v - t * rts;
            uint32_t *e = sm + t * d * rts + r;
            uint32_t run = 0;
            for (int k = 0; k < d; ++k) {
                const uint32_t c = e[k * rts];
                if (!inverse) { run = ab_addmod(run, c, q); e[k * rts] = run; }
                else { e[k * rts] = ab_submod(c, run, q); run = c; }
            }
        }
        __syncthreads();
    }
}

template <bool INV>
__global__ void __launch_bounds__(256, AB_MX_MINB) k_crt_mixed(const __grid_constant__ MxArgs args, int64_t *__restrict__ y) {
    extern __shared__ __align__(16) uint32_t smem[];
    const int limb = blockIdx.y;
    const bool wide = args.limb[limb].wide != 0;
    const int tab_words = args.tab_words[wide ? 1 : 0];
    uint32_t *tab = smem;                          // [tab_words]
    uint32_t *sm = smem + tab_words;               // [n]
    const uint32_t q = args.limb[limb].q, qinv_neg = args.limb[limb].qinv_neg;
    const int n = args.n;
    int64_t *yp = y + blockIdx.x * args.bstride + limb * args.lstride;
    {   // constant tables of this limb (L2-resident, a few KB)
        const uint4 *src = reinterpret_cast<const uint4 *>(args.limb[limb].tab);
        uint4 *dst = reinterpret_cast<uint4 *>(tab);
        for (int i = threadIdx.x; i < tab_words / 4; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    bool skip_program = false;
    if (args.fx.load_mode == AB_MX_LOAD_DIGITS) {
        // the residues of this transform are gadget digit d of a powerful-basis element (optionally gathered
        // through an index table: coeffs . embed of the tunnel), centred and reduced into this CTA's modulus
        const AbMxFused &fx = args.fx;
        const int d = (int)(blockIdx.x % fx.nd);
        const int64_t unit = blockIdx.x / fx.nd;
        if (mx_digit_is_trivial(fx, d, limb)) {
            // TrivGad digit of limb l reduced mod q_l is the element itself: its CRT is the caller's CRT-basis input
            const int4 *src = reinterpret_cast<const int4 *>(fx.triv_src + unit * fx.triv_ustride + (int64_t)limb * n);
            int4 *dst = reinterpret_cast<int4 *>(yp);
            for (int i = threadIdx.x; i < n / 2; i += blockDim.x) __stcs(dst + i, __ldcs(src + i));
            return;
        }
        mx_load_digit(fx, d, fx.gK > 1 ? unit / fx.gK : unit, fx.gK > 1 ? (int)(unit % fx.gK) : 0, n, q, sm);
    } else if (args.fx.load_mode == AB_MX_LOAD_GARNER) {
        const AbMxFused &fx = args.fx;
        const int k = fx.gar_k;
        const int64_t *src = fx.src + blockIdx.x * fx.src_ustride;
        uint32_t qmod[AB_MX_MAX_DROP];                 // q_i mod q of this CTA
#pragma unroll
        for (int i = 0; i < AB_MX_MAX_DROP; ++i) qmod[i] = i < k ? fx.gar_q[i] % q : 0u;
        const uint64_t mu = ~0ull / q;                 // Barrett constant of this CTA's modulus (floor((2^64 - 1) / q) = floor(2^64 / q), q odd > 1)
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            uint32_t v[AB_MX_MAX_DROP];
            int32_t r[AB_MX_MAX_DROP];
#pragma unroll
            for (int i = 0; i < AB_MX_MAX_DROP; ++i) v[i] = i < k ? (uint32_t)__ldcs(src - (int64_t)i * n + j) : 0u;   // drop order: last limb first
#pragma unroll
            for (int i = 0; i < AB_MX_MAX_DROP; ++i) {
                if (i >= k) break;
                r[i] = ab_center(v[i], fx.gar_q[i]);
#pragma unroll
                for (int t = i + 1; t < AB_MX_MAX_DROP; ++t) {
                    if (t >= k) break;
                    const uint32_t qt = fx.gar_q[t];
                    v[t] = ab_shoup(ab_submod(v[t], mx_reduce_signed(r[i], qt), qt), fx.gar_inv[i][t], fx.gar_inv_s[i][t], qt);
                }
            }
            uint32_t acc = mx_reduce_signed(r[k - 1], q);
            for (int i = k - 2; i >= 0; --i) acc = ab_addmod(ab_mulmod(acc, qmod[i], q, mu), mx_reduce_signed(r[i], q), q);
            sm[j] = acc;
        }
    } else if (args.fx.load_mode == AB_MX_LOAD_SRC) {
        // out of place: read the transform from another planar buffer (e.g. c2 of a degree-2 ciphertext)
        const int4 *src = reinterpret_cast<const int4 *>(args.fx.src + blockIdx.x * args.fx.src_ustride + (int64_t)(args.fx.src_limb0 + limb) * n);
        uint2 *dst = reinterpret_cast<uint2 *>(sm);
        for (int i = threadIdx.x; i < n / 2; i += blockDim.x) {
            const int4 v = __ldcs(src + i);
            dst[i] = make_uint2((uint32_t)v.x, (uint32_t)v.z);
        }
    } else if (args.estride == 1) {
        const int4 *src = reinterpret_cast<const int4 *>(yp);
        uint2 *dst = reinterpret_cast<uint2 *>(sm);
        for (int i = threadIdx.x; i < n / 2; i += blockDim.x) {
            const int4 v = __ldcs(src + i);
            dst[i] = make_uint2((uint32_t)v.x, (uint32_t)v.z);
        }
    } else {
        for (int j = threadIdx.x; j < n; j += blockDim.x) sm[j] = (uint32_t)yp[j * args.estride];
    }
    (void)skip_program;
    __syncthreads();
    if (!INV && args.fx.pre_l) mx_l_pass(sm, n, args.fx, q, false);
    if (wide) run_program<INV, true>(args, sm, tab, q, qinv_neg);
    else run_program<INV, false>(args, sm, tab, q, qinv_neg);
    if (INV && args.fx.post_linv) mx_l_pass(sm, n, args.fx, q, true);
    if (args.fx.store_mode == AB_MX_STORE_SUBSCALE) {
        // RNS rescale epilogue: out = (in - x) * q_drop^-1 mod q   (x = CRT of the lifted dropped limb)
        const int64_t *in = args.fx.sub_in + blockIdx.x * args.fx.sub_ustride + (int64_t)limb * n;
        const uint32_t sc = args.fx.scale[limb], scs = args.fx.scale_s[limb];
        const int4 *src = reinterpret_cast<const int4 *>(in);
        int4 *dst = reinterpret_cast<int4 *>(yp);
        const uint2 *xs = reinterpret_cast<const uint2 *>(sm);
        for (int i = threadIdx.x; i < n / 2; i += blockDim.x) {
            const int4 a = __ldcs(src + i);
            const uint2 x = xs[i];
            const uint32_t r0 = ab_shoup(ab_submod((uint32_t)a.x, x.x, q), sc, scs, q);
            const uint32_t r1 = ab_shoup(ab_submod((uint32_t)a.z, x.y, q), sc, scs, q);
            __stcs(dst + i, make_int4((int)r0, 0, (int)r1, 0));
        }
    } else if (args.estride == 1) {
        int4 *dst = reinterpret_cast<int4 *>(yp);
        const uint2 *src = reinterpret_cast<const uint2 *>(sm);
        for (int i = threadIdx.x; i < n / 2; i += blockDim.x) {
            const uint2 v = src[i];
            __stcs(dst + i, make_int4((int)v.x, 0, (int)v.y, 0));
        }
    } else {
        for (int j = threadIdx.x; j < n; j += blockDim.x) yp[j * args.estride] = (int64_t)sm[j];
    }
}

// ---------------------------------------------------------------------------------------------
// Fused gadget inner product: out_c = add_c + sum_{k,d} CRT(reduce(digit_{k,d}(src))) * hint_{k,d,c},  c in {0,1}.
// The core of keySwitchQuadCirc (Eval.hs:133) and of tunnel (Eval.hs:134).  One CTA owns (unit, limb) and keeps
// the residues of the digit being transformed (uint32 [n]) and the two running sums (uint2 [n], fully reduced)
// in shared memory for the whole loop: the K * nd digit transforms never touch HBM.  Per digit: load + centre +
// digit + reduce (source and gather table from L2), the forward program, then one pass x * (h0, h1) with a
// single-term Montgomery reduction each (T = x h 2^32 < 2^62, T + m q < 2^64 for q < 2^31).
// ACC_SMEM = false: the running sums live in a global scratch slice owned by this CTA (thread t always touches the
// same positions j = t, t + T, ..., so no synchronisation is needed; with one wave of CTAs resident the slices stay
// in L2) and shared memory holds only the residues: the 4 x 256-thread residency of the plain transform is kept.
template <int MAXT, int MINB, bool ACC_SMEM>
__global__ void __launch_bounds__(MAXT, MINB) k_crt_mixed_mac(const __grid_constant__ MxArgs args, const __grid_constant__ AbMxMac mac) {
    extern __shared__ __align__(16) uint32_t smem[];
    const int limb = blockIdx.y;
    const bool wide = args.limb[limb].wide != 0;
    const int tab_words = args.tab_words[wide ? 1 : 0];
    const int n = args.n, nb = (n + 3) & ~3;
    uint32_t *tab = smem;
    uint32_t *sm = smem + tab_words;
    const uint32_t q = args.limb[limb].q, qinv_neg = args.limb[limb].qinv_neg;
    const AbMxFused &fx = args.fx;
    const int parts = mac.parts > 1 ? mac.parts : 1;
    const int64_t unit = blockIdx.x / parts, slice = blockIdx.x;          // slice = unit * parts + part
    const int part = (int)(blockIdx.x - unit * parts);
    uint2 *acc = ACC_SMEM ? reinterpret_cast<uint2 *>(sm + nb)
                          : reinterpret_cast<uint2 *>(mac.acc) + (slice * args.nlimbs + limb) * (int64_t)n;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(args.limb[limb].tab);
        uint4 *dst = reinterpret_cast<uint4 *>(tab);
        for (int i = threadIdx.x; i < tab_words / 4; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    int64_t *tbuf = (mac.tbuf && fx.gbase != 0) ? mac.tbuf + (slice * args.nlimbs + limb) * (int64_t)n : nullptr;
    const uint2 *hm = reinterpret_cast<const uint2 *>(mac.hm) + (int64_t)limb * n;
    const int64_t hstride = (int64_t)args.nlimbs * n;
    bool first = true;
    const int total_digits = mac.K * fx.nd, per_part = (total_digits + parts - 1) / parts;
    const int kd0 = part * per_part, kd1 = min(total_digits, kd0 + per_part);
    {
        for (int kd = kd0; kd < kd1; ++kd) {
            const int k = kd / fx.nd, d = kd - k * fx.nd;
            const bool triv = mx_digit_is_trivial(fx, d, limb);
            if (triv) {
                const int4 *src = reinterpret_cast<const int4 *>(fx.triv_src + unit * fx.triv_ustride + (int64_t)limb * n);
                uint2 *dst = reinterpret_cast<uint2 *>(sm);
                for (int i = threadIdx.x; i < n / 2; i += blockDim.x) {
                    const int4 v = __ldcs(src + i);
                    dst[i] = make_uint2((uint32_t)v.x, (uint32_t)v.z);
                }
            } else if (tbuf) {
                if (kd == kd0 || d == 0 || fx.dig[d].limb != fx.dig[d - 1].limb) {
                    mx_gather_centered(fx, d, unit, k, n, tbuf);
                    __syncthreads();      // the digit pass reads t with a different thread-to-position map
                }
                mx_digit_from_t(fx, d, n, q, tbuf, sm);
            } else {
                mx_load_digit(fx, d, unit, k, n, q, sm);
            }
            __syncthreads();
            if (!triv) {
                if (wide) run_program<false, true>(args, sm, tab, q, qinv_neg);
                else run_program<false, false>(args, sm, tab, q, qinv_neg);
            }
            const uint2 *hp = hm + (int64_t)(k * fx.nd + d) * hstride;
            auto mac1 = [&](uint32_t x, uint32_t h0, uint32_t h1, uint32_t &a0, uint32_t &a1) {
                const uint64_t t0 = (uint64_t)x * h0, t1 = (uint64_t)x * h1;
                const uint32_t r0 = (uint32_t)((t0 + (uint64_t)((uint32_t)t0 * qinv_neg) * q) >> 32);
                const uint32_t r1 = (uint32_t)((t1 + (uint64_t)((uint32_t)t1 * qinv_neg) * q) >> 32);
                a0 = ab_addmod(ab_csub(r0, q), a0, q);
                a1 = ab_addmod(ab_csub(r1, q), a1, q);
            };
            // four consecutive positions per thread and step: 16-byte loads of the hint pairs, the running sums and the
            // residues; thread t owns the same positions for every digit (the sums need no synchronisation)
            const int T = blockDim.x, n4 = n >> 2;
            const uint4 *h4 = reinterpret_cast<const uint4 *>(hp);
            uint4 *a4 = reinterpret_cast<uint4 *>(acc);
            const uint4 *x4 = reinterpret_cast<const uint4 *>(sm);
            for (int g = threadIdx.x; g < n4; g += T) {
                const uint4 ha = __ldg(h4 + 2 * g), hb2 = __ldg(h4 + 2 * g + 1);
                uint4 aa = make_uint4(0, 0, 0, 0), ab2 = aa;
                if (!first) {
                    if (ACC_SMEM) { aa = a4[2 * g]; ab2 = a4[2 * g + 1]; }
                    else { aa = __ldcg(a4 + 2 * g); ab2 = __ldcg(a4 + 2 * g + 1); }
                }
                const uint4 x = x4[g];
                mac1(x.x, ha.x, ha.y, aa.x, aa.y);
                mac1(x.y, ha.z, ha.w, aa.z, aa.w);
                mac1(x.z, hb2.x, hb2.y, ab2.x, ab2.y);
                mac1(x.w, hb2.z, hb2.w, ab2.z, ab2.w);
                if (ACC_SMEM) { a4[2 * g] = aa; a4[2 * g + 1] = ab2; }
                else { __stcg(a4 + 2 * g, aa); __stcg(a4 + 2 * g + 1, ab2); }
            }
            for (int j = 4 * n4 + threadIdx.x; j < n; j += T) {
                const uint2 h = __ldg(hp + j);
                uint2 a = make_uint2(0, 0);
                if (!first) a = ACC_SMEM ? acc[j] : __ldcg(acc + j);
                mac1(sm[j], h.x, h.y, a.x, a.y);
                if (ACC_SMEM) acc[j] = a;
                else __stcg(acc + j, a);
            }
            first = false;
            __syncthreads();       // the next digit overwrites sm
        }
    }
    if (first && !ACC_SMEM) {       // a part that received no digit (parts does not divide the digit count): its partial sum is zero
        for (int j = threadIdx.x; j < n; j += blockDim.x) __stcg(acc + j, make_uint2(0, 0));
    }
    if (parts == 1) {
        const int64_t per = hstride;
        int64_t *o0 = mac.out + unit * mac.out_ustride + (int64_t)limb * n, *o1 = o0 + per;
        const int64_t *a0 = mac.add ? mac.add + unit * mac.add_ustride + (int64_t)limb * n : nullptr;
        const int64_t *a1 = a0 ? a0 + per : nullptr;
        const int64_t *l0 = mac.add0 ? mac.add0 + unit * mac.add0_ustride + (int64_t)limb * n : nullptr;
        for (int i = threadIdx.x; i < n / 2; i += blockDim.x) {
            // positions 2i, 2i+1 were accumulated by other threads of this CTA: the barrier after the last digit orders them
            const uint2 u = ACC_SMEM ? acc[2 * i] : __ldcg(acc + 2 * i), v = ACC_SMEM ? acc[2 * i + 1] : __ldcg(acc + 2 * i + 1);
            uint32_t r00 = u.x, r01 = v.x, r10 = u.y, r11 = v.y;      // r[c][column]
            if (a0) {
                const int4 x0 = __ldcs(reinterpret_cast<const int4 *>(a0) + i), x1 = __ldcs(reinterpret_cast<const int4 *>(a1) + i);
                r00 = ab_addmod(r00, (uint32_t)x0.x, q); r01 = ab_addmod(r01, (uint32_t)x0.z, q);
                r10 = ab_addmod(r10, (uint32_t)x1.x, q); r11 = ab_addmod(r11, (uint32_t)x1.z, q);
            }
            if (l0) {
                const int4 x0 = __ldcs(reinterpret_cast<const int4 *>(l0) + i);
                r00 = ab_addmod(r00, (uint32_t)x0.x, q); r01 = ab_addmod(r01, (uint32_t)x0.z, q);
            }
            __stcs(reinterpret_cast<int4 *>(o0) + i, make_int4((int)r00, 0, (int)r01, 0));
            __stcs(reinterpret_cast<int4 *>(o1) + i, make_int4((int)r10, 0, (int)r11, 0));
        }
    }
}

// out_c = add_c + sum over the parts of a unit of the partial sums (parts > 1 only)
__global__ void __launch_bounds__(256) k_mac_combine(const __grid_constant__ MxArgs args, const __grid_constant__ AbMxMac mac) {
    const int n = args.n, L = args.nlimbs, parts = mac.parts;
    const int64_t per = (int64_t)L * n, total = args.batch * per;
    const uint2 *acc = reinterpret_cast<const uint2 *>(mac.acc);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t unit = i / per, r = i - unit * per;
        const uint32_t q = args.limb[(int)(r / n)].q;
        uint32_t s0 = 0, s1 = 0;
        for (int pt = 0; pt < parts; ++pt) {
            const uint2 a = __ldcs(acc + (unit * parts + pt) * per + r);
            s0 = ab_addmod(s0, a.x, q); s1 = ab_addmod(s1, a.y, q);
        }
        if (mac.add) {
            s0 = ab_addmod(s0, (uint32_t)__ldcs(mac.add + unit * mac.add_ustride + r), q);
            s1 = ab_addmod(s1, (uint32_t)__ldcs(mac.add + unit * mac.add_ustride + per + r), q);
        }
        if (mac.add0) s0 = ab_addmod(s0, (uint32_t)__ldcs(mac.add0 + unit * mac.add0_ustride + r), q);
        __stcs(mac.out + unit * mac.out_ustride + r, (int64_t)s0);
        __stcs(mac.out + unit * mac.out_ustride + per + r, (int64_t)s1);
    }
}

bool ab_mixed_supported(const ab_plan *p, const AbView &v) {
    if (p->mx_ok.empty() || p->mx_fwd.empty() || p->mx_tab_fwd.empty()) return false;
    for (int l = 0; l < v.nlimbs; ++l)
        if (!p->mx_ok[v.limb0 + l]) return false;
    if (v.layout == AB_LAYOUT_PLANAR && (((uintptr_t)v.y & 15) != 0 || (p->n & 1))) return false;
    return true;
}

// CTA size: every pass hands ceil(items / T) work items to each thread and ends in a barrier, so a T that does not
// divide the item counts leaves threads idle in the last round (n = 9216, 12-point steps, two columns per item:
// 384 items -> 1.5 rounds at T = 256).  Model: time of one transform ~ sum over passes of rounds * work per item,
// throughput ~ resident CTAs / time; resident CTAs from shared memory and the 64-register budget (1024 threads).
// Measured on B200 (tools/sweep_crt.py, AB_MX_T): within 3 % of the best of {128, 160, 192, 224, 256} on every example ring.
static int pick_threads_among(const std::vector<AbMxStep> &steps, int n, std::initializer_list<int> cand) {
    int best_t = *cand.begin();
    double best = 1e300;
    for (int T : cand) {
        double time = 6.0 * ((n / 2 + T - 1) / T) + 20.0 * ((n + T - 1) / T);      // result store; digit load + multiply-accumulate
        for (const AbMxStep &s : steps) {
            const int items = s.nwork / s.vec, rounds = (items + T - 1) / T;
            const double w = s.kind == 0 ? (double)s.span * s.vec * (s.span + 6) : (double)s.span * s.vec * (3.5 * s.g + 2);
            time += rounds * w;
        }
        // idle threads still hold registers: prefer the T with the least time, then the smaller footprint
        if (time < best * 0.9999 || (time < best * 1.0001 && T > best_t)) { best = time; best_t = T; }
    }
    return best_t;
}
static int pick_threads(const std::vector<AbMxStep> &steps, int n, size_t smem) {
    int best_t = 256;
    double best = 0;
    for (int T : {128, 160, 192, 256}) {
        const int ctas = (int)std::min<size_t>((size_t)(1024 / T), (size_t)(227 * 1024) / (smem + 1024));
        if (ctas < 1) continue;
        double time = 6.0 * ((n / 2 + T - 1) / T);                 // load + store phases
        for (const AbMxStep &s : steps) {
            const int items = s.nwork / s.vec, rounds = (items + T - 1) / T;
            const double w = s.kind == 0 ? (double)s.span * s.vec * (s.span + 6) : (double)s.span * s.vec * (3.5 * s.g + 2);
            time += rounds * w;
        }
        const double tput = ctas / time;
        if (tput > best * 1.0001) { best = tput; best_t = T; }
    }
    return best_t;
}

static uint32_t neg_inv32(uint32_t q) {   // -q^-1 mod 2^32 (Newton)
    uint32_t x = q;                        // q odd: x = q is an inverse mod 8
    for (int i = 0; i < 5; ++i) x *= 2u - q * x;
    return 0u - x;
}

void ab_launch_crt_mixed(const ab_plan *p, const AbView &v, bool inverse, cudaStream_t st) {
    ab_launch_crt_mixed_fused(p, v, inverse, nullptr, st);
}

// `v` describes the DESTINATION (planar): v.batch = number of transforms per limb (units * nd for digit loads)
void ab_launch_crt_mixed_fused(const ab_plan *p, const AbView &v, bool inverse, const AbMxFused *fx, cudaStream_t st) {
    AB_REQUIRE(ab_mixed_supported(p, v), AB_ERR_UNSUPPORTED, "mixed kernel: plan not supported");
    AB_REQUIRE(fx == nullptr || v.layout == AB_LAYOUT_PLANAR, AB_ERR_UNSUPPORTED, "fused modes need the planar layout");
    if (v.batch == 0) return;
    MxArgs args{};
    if (fx) args.fx = *fx;
    if (args.fx.pre_l || args.fx.post_linv) {
        args.fx.n_odd = 0;
        for (const AbAxis &ax : p->axes)
            if (ax.p != 2) { args.fx.odd[args.fx.n_odd].d = ax.p - 1; args.fx.odd[args.fx.n_odd].rts = (int32_t)ax.rts; ++args.fx.n_odd; }
    }
    const auto &steps = inverse ? p->mx_inv : p->mx_fwd;
    args.nsteps = (int32_t)steps.size();
    for (size_t i = 0; i < steps.size(); ++i) args.steps[i] = steps[i];
    args.n = (int32_t)p->n;
    args.tab_words[0] = inverse ? p->mx_words_inv[0] : p->mx_words_fwd[0];
    args.tab_words[1] = inverse ? p->mx_words_inv[1] : p->mx_words_fwd[1];
    int max_words = 0;
    args.nlimbs = v.nlimbs;
    args.batch = v.batch;
    if (v.layout == AB_LAYOUT_PLANAR) { args.estride = 1; args.lstride = p->n; args.bstride = p->n * v.nlimbs; }
    else { args.estride = v.nlimbs; args.lstride = 1; args.bstride = p->n * v.nlimbs; }
    if (args.fx.dst_ustride) args.bstride = args.fx.dst_ustride;
    for (int l = 0; l < v.nlimbs; ++l) {
        const AbLimbDev &ld = p->limb[v.limb0 + l];
        args.limb[l].q = ld.q;
        args.limb[l].qinv_neg = neg_inv32(ld.q);
        args.limb[l].tab = inverse ? p->mx_tab_inv[v.limb0 + l] : p->mx_tab_fwd[v.limb0 + l];
        args.limb[l].wide = p->mx_ok[v.limb0 + l] == 2;
        max_words = std::max(max_words, args.tab_words[args.limb[l].wide]);
    }
    const size_t smem = ((size_t)max_words + (size_t)p->n) * sizeof(uint32_t);
    AB_REQUIRE(smem <= 227 * 1024, AB_ERR_UNSUPPORTED, "ring too large for the shared-memory transform");
    static thread_local bool ready = false;
    if (!ready) {
        AB_CHECK_CUDA(cudaFuncSetAttribute(k_crt_mixed<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        AB_CHECK_CUDA(cudaFuncSetAttribute(k_crt_mixed<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        ready = true;
    }
    // grid.x is limited to 2^31-1, grid.y = limbs
    const dim3 grid((unsigned)v.batch, (unsigned)v.nlimbs);
    static const int force_t = getenv("AB_MX_T") ? atoi(getenv("AB_MX_T")) : 0;      // experiments only
    const int T = force_t ? force_t : pick_threads(steps, (int)p->n, smem);
    if (inverse) k_crt_mixed<true><<<grid, T, smem, st>>>(args, v.y);
    else k_crt_mixed<false><<<grid, T, smem, st>>>(args, v.y);
    ab_count_launch();
    ab_check_launch("crt_mixed");
}

static void fill_mx_args(const ab_plan *p, MxArgs &args, int32_t limb0, int32_t nlimbs, bool inverse, int &max_words) {
    const auto &steps = inverse ? p->mx_inv : p->mx_fwd;
    args.nsteps = (int32_t)steps.size();
    for (size_t i = 0; i < steps.size(); ++i) args.steps[i] = steps[i];
    args.n = (int32_t)p->n;
    args.tab_words[0] = inverse ? p->mx_words_inv[0] : p->mx_words_fwd[0];
    args.tab_words[1] = inverse ? p->mx_words_inv[1] : p->mx_words_fwd[1];
    args.nlimbs = nlimbs;
    max_words = 0;
    for (int l = 0; l < nlimbs; ++l) {
        const AbLimbDev &ld = p->limb[limb0 + l];
        args.limb[l].q = ld.q;
        args.limb[l].qinv_neg = neg_inv32(ld.q);
        args.limb[l].tab = inverse ? p->mx_tab_inv[limb0 + l] : p->mx_tab_fwd[limb0 + l];
        args.limb[l].wide = p->mx_ok[limb0 + l] == 2;
        max_words = std::max(max_words, args.tab_words[args.limb[l].wide]);
    }
}

static size_t mac_smem(const ab_plan *p, int max_words, bool acc_smem) {
    const size_t nb = ((size_t)p->n + 3) & ~(size_t)3;
    return ((size_t)max_words + nb + (acc_smem ? 2 * (size_t)p->n : 0)) * sizeof(uint32_t);
}
// the running sums ride in shared memory only where four CTAs still fit an SM; otherwise in the caller's scratch
static bool mac_acc_in_smem(const ab_plan *p, int max_words) { return 4 * (mac_smem(p, max_words, true) + 1024) <= 228 * 1024; }
bool ab_mixed_mac_supported(const ab_plan *p, int32_t nlimbs) {
    if (ab_pow2_supported(p) || (p->n & 1)) return false;
    return ab_mixed_supported(p, AbView{nullptr, 1, 0, nlimbs, AB_LAYOUT_INTERLEAVED});
}
// a CTA of the fused inner product runs for milliseconds: when the batch alone gives fewer than about four waves of CTAs, cut the
// digit list of a unit into parts (each keeps at least eight digit transforms, so table loads and gathers stay amortised).
// Measured on the Tunnel.hs chain: 512 ciphertexts 34.7k -> 40.2k/s; at 4096 ciphertexts splitting does not pay (left at 1 part).
int32_t ab_mixed_mac_parts(const ab_plan *p, int64_t units, int32_t nlimbs, int32_t total_digits) {
    static const int force = getenv("AB_MX_MAC_PARTS") ? atoi(getenv("AB_MX_MAC_PARTS")) : 0;      // experiments only
    if (force > 0) return std::min(force, total_digits);
    if (units <= 0) return 1;
    const int64_t slots = (int64_t)p->sm_count * 5, want = 4 * slots;
    int64_t parts = (want + units * nlimbs - 1) / (units * nlimbs);
    parts = std::min<int64_t>(parts, total_digits / 8);
    return (int32_t)std::max<int64_t>(1, parts);
}
int64_t ab_mixed_mac_scratch_words(const ab_plan *p, int64_t units, int32_t nlimbs) {
    const int words = std::max(p->mx_words_fwd[0], p->mx_words_fwd[1]);
    return mac_acc_in_smem(p, words) ? 0 : units * nlimbs * p->n;      // uint2 per position = one int64 word (per part when parts > 1)
}

template <bool ACC_SMEM>
static void launch_mac(const MxArgs &args, const AbMxMac &mac, dim3 grid, int T, size_t smem, cudaStream_t st) {
    static thread_local bool ready = false;
    if (!ready) {
        AB_CHECK_CUDA(cudaFuncSetAttribute(k_crt_mixed_mac<256, AB_MX_MINB, ACC_SMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        ready = true;
    }
    k_crt_mixed_mac<256, AB_MX_MINB, ACC_SMEM><<<grid, T, smem, st>>>(args, mac);
}

void ab_launch_crt_mixed_mac(const ab_plan *p, int64_t units, int32_t nlimbs, const AbMxFused &fx, const AbMxMac &mac, cudaStream_t st) {
    AB_REQUIRE(ab_mixed_mac_supported(p, nlimbs), AB_ERR_UNSUPPORTED, "fused gadget inner product: plan not supported");
    AB_REQUIRE(fx.load_mode == AB_MX_LOAD_DIGITS && fx.nd >= 1 && mac.K >= 1 && mac.hm && mac.out, AB_ERR_ARG, "fused gadget inner product: bad arguments");
    if (units == 0) return;
    MxArgs args{};
    args.fx = fx;
    int max_words = 0;
    fill_mx_args(p, args, 0, nlimbs, false, max_words);
    args.batch = units;
    args.estride = 1; args.lstride = p->n; args.bstride = p->n * nlimbs;
    const int parts = mac.parts > 1 ? mac.parts : 1;
    const bool acc_smem = parts == 1 && mac_acc_in_smem(p, std::max(p->mx_words_fwd[0], p->mx_words_fwd[1]));
    AB_REQUIRE(acc_smem || mac.acc, AB_ERR_ARG, "fused gadget inner product: scratch missing");
    const size_t smem = mac_smem(p, max_words, acc_smem);
    static const int force_t = getenv("AB_MX_MAC_T") ? atoi(getenv("AB_MX_MAC_T")) : 0;      // experiments only
    const int T = force_t ? force_t : pick_threads(p->mx_fwd, (int)p->n, smem);
    const dim3 grid((unsigned)(units * parts), (unsigned)nlimbs);
    if (acc_smem) launch_mac<true>(args, mac, grid, T, smem, st);
    else launch_mac<false>(args, mac, grid, T, smem, st);
    ab_count_launch();
    ab_check_launch("crt_mixed_mac");
    if (parts > 1) {
        const int64_t work = units * nlimbs * p->n;
        const unsigned g = (unsigned)std::max<int64_t>(1, std::min<int64_t>((work + 255) / 256, (int64_t)p->sm_count * 16));
        k_mac_combine<<<g, 256, 0, st>>>(args, mac);
        ab_count_launch();
        ab_check_launch("mac_combine");
    }
}
