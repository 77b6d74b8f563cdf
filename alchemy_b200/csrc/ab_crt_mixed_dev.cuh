// ab_crt_mixed_dev.cuh -- device code of the mixed-radix CRT program (dense Montgomery steps, radix-2 groups, digit
// loaders, L / L^-1 passes) and the host helpers shared by its two translation units: ab_crt_mixed.cu (the transform
// kernels) and ab_crt_mixed_mac.cu (the fused gadget inner product).  Two files so that they compile in parallel.
#pragma once
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
        int32_t wide;              // table layout of this limb (ab_mixed_plan.cpp)
        int32_t lazy;              // fused inner product: how many products x * h (both below q) one 64-bit sum may hold before its Montgomery step
    } limb[AB_MAX_LIMBS];
    int32_t nsteps, n, nlimbs;
    int32_t interleave, pad_il;    // grid order: 0 = limb-major (default), 1 = limbs interleaved (AB_MX_INTERLEAVE=1)
    int32_t tab_words[2];          // narrow, wide
    int64_t batch;
    int64_t estride, lstride, bstride;   // int64 strides of coefficient, limb, ring element
    AbMxFused fx;                        // fused load / store (ab_internal.h); all zero = plain in-place transform
};

#ifndef AB_MX_MINB
#define AB_MX_MINB 5          // resident 256-thread CTAs per SM the register budget of the transform kernels is tuned for (48 registers;
                              // measured against 4 / 64 registers: +1 ... +5 % on the example rings, inverse transforms most)
#endif
#ifndef AB_MX_MAC_MINB
#define AB_MX_MAC_MINB 4      // the fused inner product keeps 64 registers (an 80-register build measured slower, fewer registers spill)
#endif
#ifndef AB_MX_REDC_HI
#define AB_MX_REDC_HI 0       // 1: Montgomery step as hi + umulhi(m, q) + (lo != 0) (an IMAD.HI and two ALU ops) instead of a 64-bit IMAD.WIDE
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
// WIDE  (q < 2^31): inputs and constants centred, T = sum x_c M_c as a signed 64-bit sum of products below q^2 / 4 (at most
// seven per sum; the 10- to 12-point steps use two sums), one signed Montgomery reduction per sum.
// CONTIG (V == 1, element stride 1: the steps of the last tensor axis): the P residues of an item are contiguous, so they
// are moved with 16- or 8-byte shared-memory accesses.  With 4-byte accesses the lanes of a warp sit P words apart and,
// for P = 12, four of them share every bank (measured: 70 % of all excess shared-memory wavefronts of the transform);
// lanes reading 16 bytes at a 48-byte pitch cover the 32 banks exactly once per quarter warp.
// (T + m q) / 2^32 for T < 2^64 - 2^32 q, m = lo(T) * (-q^-1): the low words cancel to 0 mod 2^32 and carry out exactly when lo(T) != 0
__device__ __forceinline__ uint32_t mx_redc_u(uint64_t T, uint32_t q, uint32_t qinv_neg) {
    const uint32_t lo = (uint32_t)T, mq = lo * qinv_neg;
#if AB_MX_REDC_HI
    return (uint32_t)(T >> 32) + __umulhi(mq, q) + (lo != 0u ? 1u : 0u);
#else
    return (uint32_t)((T + (uint64_t)mq * q) >> 32);
#endif
}
// signed helpers of the WIDE path (2^30 <= q < 2^31): centred inputs and constants, signed 64-bit sums
__device__ __forceinline__ int32_t mx_centre(uint32_t x, uint32_t q, uint32_t hq) { return (int32_t)(x > hq ? x - q : x); }      // [0,q) -> [-hq, hq]
// signed Montgomery: T 2^-32 mod q for a signed sum T, result in [0,q).  m = lo(T) q^-1 mod 2^32 taken UNSIGNED (q^-1 = -qinv_neg): the low
// words of T and m q agree, so (T - m q) / 2^32 = hi(T) - hi(m q) exactly, a value in (-|T|/2^32 - q, |T|/2^32].  Adding 2q makes it
// positive and the plan's bound check keeps it below 2^32 (|T|/2^32 + 2q < 2^32), so the 32-bit arithmetic below is exact: one IMAD,
// one IMAD.HI, one IADD3 and two conditional subtractions (a signed m needs sign tests on the result: four more instructions).
__device__ __forceinline__ uint32_t mx_redc_s(int64_t T, uint32_t q, uint32_t qinv_neg) {
    const uint32_t m = (uint32_t)T * (0u - qinv_neg);
    const uint32_t t = (uint32_t)(T >> 32) - __umulhi(m, q) + 2u * q;      // in (0, 3q) as an integer
    return ab_csub(ab_csub(t, 2u * q), q);
}

template <int P, int V, bool WIDE, bool CONTIG = false>
__device__ __forceinline__ void matvec_item(uint32_t *sm, const uint32_t *__restrict__ tab, const AbMxStep &s, int v, uint32_t q, uint32_t qinv_neg) {
    static_assert(!CONTIG || V == 1, "contiguous items have one column");
    constexpr int PP = (P + 3) & ~3;
    constexpr int NA = (WIDE && P > 7) ? 2 : 1, PH = (P + NA - 1) / NA;     // WIDE: at most seven signed products per 64-bit sum
    int tb, off;
    item_coords<V>(s, v, tb, off);
    const int es = s.estride;
    uint32_t *e = sm + tb * (P * es) + off;
    const uint4 *M = reinterpret_cast<const uint4 *>(tab + (WIDE ? s.tab_off_w : s.tab_off) + item_row(s, tb, off) * (P * PP));
    uint32_t x[P][V];
    if (CONTIG) {
        if (P % 4 == 0) {
            mx_static_for<(P % 4 == 0 ? P / 4 : 0)>([&](auto K) {
                constexpr int k = decltype(K)::value;
                const uint4 t = reinterpret_cast<const uint4 *>(e)[k];
                x[(4 * k) % P][0] = t.x; x[(4 * k + 1) % P][0] = t.y; x[(4 * k + 2) % P][0] = t.z; x[(4 * k + 3) % P][0] = t.w;
            });
        } else if (P % 2 == 0) {
            mx_static_for<(P % 2 == 0 ? P / 2 : 0)>([&](auto K) {
                constexpr int k = decltype(K)::value;
                const uint2 t = reinterpret_cast<const uint2 *>(e)[k];
                x[(2 * k) % P][0] = t.x; x[(2 * k + 1) % P][0] = t.y;
            });
        } else {
            mx_static_for<P>([&](auto C) { x[decltype(C)::value][0] = e[decltype(C)::value]; });
        }
    } else {
        mx_static_for<P>([&](auto C) { ld_cols<V>(e + decltype(C)::value * es, x[decltype(C)::value]); });
    }
    uint32_t oo[CONTIG ? P : 1];
    if (WIDE) {
        const uint32_t hq = (q - 1) >> 1;
        mx_static_for<P>([&](auto C) {
            constexpr int c = decltype(C)::value;
#pragma unroll
            for (int k = 0; k < V; ++k) x[c][k] = (uint32_t)mx_centre(x[c][k], q, hq);
        });
    }
    mx_static_for<P>([&](auto R) {
        constexpr int r = decltype(R)::value;
        uint64_t acc[V], acc2[V];
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] = acc2[k] = 0;
        mx_static_for<PP / 4>([&](auto K) {
            constexpr int k4 = decltype(K)::value;
            const uint4 m4 = M[r * (PP / 4) + k4];
            const uint32_t mm[4] = {m4.x, m4.y, m4.z, m4.w};
            mx_static_for<4>([&](auto J) {
                constexpr int c = 4 * k4 + decltype(J)::value;
                if (c < P) {
#pragma unroll
                    for (int k = 0; k < V; ++k) {
                        if (WIDE) {
                            const int64_t t = (int64_t)(int32_t)x[c % P][k] * (int32_t)mm[decltype(J)::value];
                            if (NA == 2 && c >= PH) acc2[k] += (uint64_t)t;
                            else acc[k] += (uint64_t)t;
                        } else if (AB_MX_SPLIT && (c & 1)) acc2[k] += (uint64_t)x[c % P][k] * mm[decltype(J)::value];
                        else acc[k] += (uint64_t)x[c % P][k] * mm[decltype(J)::value];
                    }
                }
            });
        });
        uint32_t o[V];
#pragma unroll
        for (int k = 0; k < V; ++k) {
            if (WIDE) {
                o[k] = mx_redc_s((int64_t)acc[k], q, qinv_neg);
                if (NA == 2) o[k] = ab_addmod(o[k], mx_redc_s((int64_t)acc2[k], q, qinv_neg), q);
            } else {
                if (AB_MX_SPLIT) acc[k] += acc2[k];
                uint32_t t = mx_redc_u(acc[k], q, qinv_neg);
                if (P > 4) t = ab_csub(t, 2 * q);
                o[k] = ab_csub(t, q);
            }
        }
        if (CONTIG) oo[CONTIG ? r : 0] = o[0];
        else st_cols<V>(e + r * es, o);      // all inputs are already in registers: in place is safe
    });
    if (CONTIG) {
        if (P % 4 == 0) {
            mx_static_for<(P % 4 == 0 ? P / 4 : 0)>([&](auto K) {
                constexpr int k = decltype(K)::value;
                reinterpret_cast<uint4 *>(e)[k] = make_uint4(oo[(4 * k) % P], oo[(4 * k + 1) % P], oo[(4 * k + 2) % P], oo[(4 * k + 3) % P]);
            });
        } else if (P % 2 == 0) {
            mx_static_for<(P % 2 == 0 ? P / 2 : 0)>([&](auto K) {
                constexpr int k = decltype(K)::value;
                reinterpret_cast<uint2 *>(e)[k] = make_uint2(oo[(2 * k) % P], oo[(2 * k + 1) % P]);
            });
        } else {
            mx_static_for<P>([&](auto C) { e[decltype(C)::value] = oo[decltype(C)::value % (CONTIG ? P : 1)]; });
        }
    }
}

// ---- CRT_p of a prime axis through its conjugate symmetry (step kind 2; tables: ab_mixed_plan.cpp push_pair_tables) ----------------
// Forward: u_j = x_j + x_(p-j), v_j = x_j - x_(p-j) (x_(p-1) = 0), out_i / out_(p-i) = x_0 + sum_j u_j C_ij +/- sum_j v_j S_ij.
// Inverse: U_k = X_k + X_(p-k), V_k = X_k - X_(p-k), y_j / y_(p-j) = sum_k U_k C'_jk -/+ sum_k V_k S'_jk, y_0 = sum_k U_k / p,
// x_j = y_j - y_(p-1).  P = p - 1 inputs and outputs, h = P / 2 products per 64-bit sum, two sums per pair of outputs: half the
// IMAD.WIDE of the dense step.  narrow: the sums u, v stay unreduced (below 2q), T < P q^2 as in the dense step; WIDE: u, v reduced
// and centred, signed sums of h <= 6 products below q^2 / 4.
template <int H, int V, bool WIDE>
__device__ __forceinline__ void pair_dot(const uint4 *__restrict__ row, const uint32_t (*xa)[V], const uint32_t (*xb)[V], uint32_t q, uint32_t qinv_neg,
                                         uint32_t (&a)[V], uint32_t (&b)[V]) {
    // row: [A_1..A_H, B_1..B_H, pad]; xa[j], xb[j] (j < H): the factors of the A and B constants (WIDE: centred, as int32)
    constexpr int PP = (2 * H + 3) & ~3;
    uint64_t sa[V], sb[V];
#pragma unroll
    for (int k = 0; k < V; ++k) sa[k] = sb[k] = 0;
    mx_static_for<PP / 4>([&](auto K) {
        constexpr int k4 = decltype(K)::value;
        const uint4 t = row[k4];
        const uint32_t c[4] = {t.x, t.y, t.z, t.w};
        mx_static_for<4>([&](auto J) {
            constexpr int idx = 4 * k4 + decltype(J)::value;
            if (idx < 2 * H) {
#pragma unroll
                for (int k = 0; k < V; ++k) {
                    const uint32_t f = idx < H ? xa[idx % H][k] : xb[idx % H][k];
                    const uint64_t pr = WIDE ? (uint64_t)((int64_t)(int32_t)f * (int32_t)c[decltype(J)::value]) : (uint64_t)f * c[decltype(J)::value];
                    if (idx < H) sa[k] += pr;
                    else sb[k] += pr;
                }
            }
        });
    });
#pragma unroll
    for (int k = 0; k < V; ++k) {
        if (WIDE) {
            a[k] = mx_redc_s((int64_t)sa[k], q, qinv_neg);
            b[k] = mx_redc_s((int64_t)sb[k], q, qinv_neg);
        } else {
            uint32_t ta = mx_redc_u(sa[k], q, qinv_neg), tb = mx_redc_u(sb[k], q, qinv_neg);
            if (H > 2) { ta = ab_csub(ta, 2 * q); tb = ab_csub(tb, 2 * q); }
            a[k] = ab_csub(ta, q);
            b[k] = ab_csub(tb, q);
        }
    }
}
template <int P, int V, bool WIDE, bool INVS, bool CONTIG = false>
__device__ __forceinline__ void pair_item(uint32_t *sm, const uint32_t *__restrict__ tab, const AbMxStep &s, int v, uint32_t q, uint32_t qinv_neg) {
    static_assert(P % 2 == 0 && P >= 4, "pair steps are for odd primes p >= 5");
    static_assert(!CONTIG || V == 1, "contiguous items have one column");
    constexpr int H = P / 2, PP = (P + 3) & ~3, RW = PP;
    int tb, off;
    item_coords<V>(s, v, tb, off);
    const int es = s.estride;
    uint32_t *e = sm + tb * (P * es) + off;
    const uint4 *M = reinterpret_cast<const uint4 *>(tab + (WIDE ? s.tab_off_w : s.tab_off));
    const uint32_t hq = (q - 1) >> 1;
    // the sums u = a + b and differences v = a - b that multiply the constants.  narrow: left unreduced (below 2q, T < P q^2);
    // WIDE: reduced and centred (|.| <= (q-1)/2), as the signed sums need
    auto usum = [&](uint32_t a, uint32_t b) { return WIDE ? (uint32_t)mx_centre(ab_addmod(a, b, q), q, hq) : a + b; };
    auto udif = [&](uint32_t a, uint32_t b) { return WIDE ? (uint32_t)mx_centre(ab_submod(a, b, q), q, hq) : a - b + q; };
    uint32_t x[P][V];
    if (CONTIG) {
        if (P % 4 == 0) {
            mx_static_for<(P % 4 == 0 ? P / 4 : 0)>([&](auto K) {
                constexpr int k = decltype(K)::value;
                const uint4 t = reinterpret_cast<const uint4 *>(e)[k];
                x[(4 * k) % P][0] = t.x; x[(4 * k + 1) % P][0] = t.y; x[(4 * k + 2) % P][0] = t.z; x[(4 * k + 3) % P][0] = t.w;
            });
        } else {
            mx_static_for<P / 2>([&](auto K) {
                constexpr int k = decltype(K)::value;
                const uint2 t = reinterpret_cast<const uint2 *>(e)[k];
                x[(2 * k) % P][0] = t.x; x[(2 * k + 1) % P][0] = t.y;
            });
        }
    } else {
        mx_static_for<P>([&](auto C) { ld_cols<V>(e + decltype(C)::value * es, x[decltype(C)::value]); });
    }
    uint32_t oo[CONTIG ? P : 1];
    auto put = [&](int r, const uint32_t (&o)[V]) {
        if (CONTIG) oo[CONTIG ? r : 0] = o[0];
        else st_cols<V>(e + r * es, o);          // all inputs are already in registers: in place is safe
    };
    if (!INVS) {
        // inputs x_0 .. x_(p-2); pairs (j, p-j) = positions (j, P+1-j) for j = 2..H; j = 1 has no partner (u_1 = v_1 = x_1)
        uint32_t xa[H][V], xb[H][V];
#pragma unroll
        for (int k = 0; k < V; ++k) {
            xa[0][k] = xb[0][k] = WIDE ? (uint32_t)mx_centre(x[1][k], q, hq) : x[1][k];
            mx_static_for<H - 1>([&](auto J) {
                constexpr int j = decltype(J)::value + 2;
                const uint32_t a = x[j][k], b = x[(P + 1 - j) % P][k];
                xa[j - 1][k] = usum(a, b);
                xb[j - 1][k] = udif(a, b);
            });
        }
        mx_static_for<H>([&](auto I) {
            constexpr int i = decltype(I)::value + 1;
            uint32_t a[V], b[V], o0[V], o1[V];
            pair_dot<H, V, WIDE>(M + (i - 1) * (RW / 4), xa, xb, q, qinv_neg, a, b);
#pragma unroll
            for (int k = 0; k < V; ++k) {
                const uint32_t t = ab_addmod(a[k], x[0][k], q);
                o0[k] = ab_addmod(t, b[k], q);
                o1[k] = ab_submod(t, b[k], q);
            }
            put(i - 1, o0);
            put(P - i, o1);
        });
    } else {
        // inputs X_1 .. X_(p-1) at positions 0 .. P-1; pairs (k, p-k) = positions (k-1, P-k)
        uint32_t xa[H][V], xb[H][V];
#pragma unroll
        for (int k = 0; k < V; ++k)
            mx_static_for<H>([&](auto J) {
                constexpr int kk = decltype(J)::value + 1;
                const uint32_t a = x[kk - 1][k], b = x[P - kk][k];
                xa[kk - 1][k] = usum(a, b);
                xb[kk - 1][k] = udif(a, b);
            });
        uint32_t y0[V], dummy[V], c[V];
        pair_dot<H, V, WIDE>(M, xa, xb, q, qinv_neg, y0, dummy);      // row 0: (1/p, ..., 1/p | 0, ..., 0)
        mx_static_for<H>([&](auto I) {
            constexpr int j = decltype(I)::value + 1;
            uint32_t a[V], b[V], o0[V], o1[V];
            pair_dot<H, V, WIDE>(M + j * (RW / 4), xa, xb, q, qinv_neg, a, b);
#pragma unroll
            for (int k = 0; k < V; ++k) {
                const uint32_t ym = ab_submod(a[k], b[k], q), yp = ab_addmod(a[k], b[k], q);      // y_j, y_(p-j)
                if (j == 1) { c[k] = yp; o0[k] = ab_submod(y0[k], c[k], q); }                    // y_(p-1); x_0 = y_0 - y_(p-1)
                o1[k] = ab_submod(ym, c[k], q);                                                   // x_j
                if (j > 1) o0[k] = ab_submod(yp, c[k], q);                                        // x_(p-j) at position P+1-j
            }
            if (j == 1) put(0, o0);
            else put(P + 1 - j, o0);
            put(j, o1);
        });
    }
    if (CONTIG) {
        if (P % 4 == 0) {
            mx_static_for<(P % 4 == 0 ? P / 4 : 0)>([&](auto K) {
                constexpr int k = decltype(K)::value;
                reinterpret_cast<uint4 *>(e)[k] = make_uint4(oo[(4 * k) % P], oo[(4 * k + 1) % P], oo[(4 * k + 2) % P], oo[(4 * k + 3) % P]);
            });
        } else {
            mx_static_for<P / 2>([&](auto K) {
                constexpr int k = decltype(K)::value;
                reinterpret_cast<uint2 *>(e)[k] = make_uint2(oo[(2 * k) % P], oo[(2 * k + 1) % P]);
            });
        }
    }
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
                            // q >= 2^30: values stay below 2q between the stages of a group (2q < 2^32), one final subtraction
                            const uint32_t a = t == 0 ? x[c][k] : ab_csub(x[c][k], q), b = ab_shoup(x[d][k], w.x, w.y, q);
                            x[c][k] = a + b;
                            x[d][k] = a - b + q;
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
            if (LAZY || !INV) t = ab_csub(t, q);
            o[k] = t;
        }
        st_cols<V>(e + c * es, o);
    });
}

// spans for which the contiguous path is compiled: the p-1 point steps of a last axis p in {5, 7, 11, 13}; the wide (two
// constant sets) variants of the long ones would need x, its high halves and the outputs live at once and spill
__host__ __device__ constexpr bool mx_contig_span(int P, bool wide) { return (P == 4 || P == 6) || (!wide && (P == 10 || P == 12)); }

template <bool INV, bool WIDE>
__device__ __forceinline__ void run_program(const MxArgs &args, uint32_t *sm, const uint32_t *tab, uint32_t q, uint32_t qinv_neg, int first_step = 0) {
    const bool contig_ok = ((uintptr_t)sm & 15) == 0;     // vector accesses need the residue buffer 16-byte aligned
    for (int si = first_step; si < args.nsteps; ++si) {
        const AbMxStep &s = args.steps[si];
        const int nw = s.nwork / s.vec;
        if (s.kind == 2) {
#define AB_PR(P)                                                                                                    \
    case P:                                                                                                         \
        if (s.vec == 2) for (int v = threadIdx.x; v < nw; v += blockDim.x) pair_item<P, 2, WIDE, INV>(sm, tab, s, v, q, qinv_neg); \
        else if (s.estride == 1 && contig_ok)                                                                       \
            for (int v = threadIdx.x; v < nw; v += blockDim.x) pair_item<P, 1, WIDE, INV, true>(sm, tab, s, v, q, qinv_neg); \
        else for (int v = threadIdx.x; v < nw; v += blockDim.x) pair_item<P, 1, WIDE, INV>(sm, tab, s, v, q, qinv_neg);      \
        break;
            switch (s.span) {
                AB_PR(4) AB_PR(6) AB_PR(10) AB_PR(12)
                default: break;
            }
#undef AB_PR
        } else if (s.kind == 0) {
#define AB_MV(P)                                                                                                    \
    case P:                                                                                                         \
        if (s.vec == 2) for (int v = threadIdx.x; v < nw; v += blockDim.x) matvec_item<P, 2, WIDE>(sm, tab, s, v, q, qinv_neg); \
        else if (mx_contig_span(P, WIDE) && s.estride == 1 && contig_ok)                                            \
            for (int v = threadIdx.x; v < nw; v += blockDim.x) matvec_item<P, 1, WIDE, mx_contig_span(P, WIDE)>(sm, tab, s, v, q, qinv_neg); \
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
// with dependent loads (index table, then coefficient), issued four at a time -- and parks the gathered residues as
// uint32 in a scratch slice owned by this CTA (half the L2 footprint of an int64 t); phase B forms t = centred + H and cuts digit d out of it with shifts (power-of-two base) or
// divisions and reduces it into this CTA's modulus.  Thread t touches the same positions in both phases.
__device__ __forceinline__ void mx_gather_centered(const AbMxFused &fx, int d0, int64_t su, int krow, int n, uint32_t *__restrict__ tbuf) {
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
            if (j < n) __stcg(tbuf + j, v[u]);
        }
    }
}
__device__ __forceinline__ void mx_digit_from_t(const AbMxFused &fx, int d, int n, uint32_t q, const uint32_t *__restrict__ tbuf, uint32_t *sm) {
    const uint32_t qs = fx.dig[d].q_src;
    const int64_t H = fx.dig[d].H;
    const int lastd = fx.dig[d].last, gb = fx.gbase, hb = fx.gbase / 2;
    const uint32_t pw = fx.dig[d].pw;
    const int pw_shift = (gb > 0 && (gb & (gb - 1)) == 0) ? 31 - __clz((int)pw) : -1;
    const int T = blockDim.x;
    if (pw_shift >= 0 && !lastd && pw_shift + (31 - __clz(gb)) <= 32 && (uint32_t)hb < q) {
        // power-of-two base, inner digit: the wanted bits of floor(t / 2^s) sit in the low word of t (two's complement),
        // and |digit| <= b/2 < q needs no reduction
        const uint32_t mask = (uint32_t)gb - 1u;
        const uint32_t Hlo = (uint32_t)H;
        auto dig = [&](uint32_t v) {
            const uint32_t lo = (uint32_t)ab_center(v, qs) + Hlo;           // low word of t = centred + H
            const int32_t c = (int32_t)((lo >> pw_shift) & mask) - hb;
            return c < 0 ? q + (uint32_t)c : (uint32_t)c;
        };
        // four consecutive positions per thread: one 16-byte load of the gathered residues, one 16-byte shared store
        const int n4 = n >> 2;
        const uint4 *t4 = reinterpret_cast<const uint4 *>(tbuf);
        uint4 *s4 = reinterpret_cast<uint4 *>(sm);
        for (int g = threadIdx.x; g < n4; g += T) {
            const uint4 a = __ldcg(t4 + g);
            s4[g] = make_uint4(dig(a.x), dig(a.y), dig(a.z), dig(a.w));
        }
        for (int j = 4 * n4 + threadIdx.x; j < n; j += T) sm[j] = dig(__ldcg(tbuf + j));
        return;
    }
    for (int j0 = threadIdx.x; j0 < n; j0 += 4 * T) {
        int64_t t[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) t[u] = (j0 + u * T < n) ? (int64_t)ab_center(__ldcg(tbuf + j0 + u * T), qs) + H : 0;
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

// Binary gadget digit d (base 2, not the last digit) of the gathered residues, and the FIRST forward step on it, in one pass:
// the digit of position j is bit d of t = centred(v_j) + H minus one, i.e. -1 or 0, so the P-point step of an item is minus the
// sum of the matrix columns whose digit is -1 -- two table lookups per output (ab_internal.h mx_bits_lut) instead of P
// multiply-adds and a reduction, and the digit never goes through shared memory.  Items are P consecutive positions.
template <int P>
__device__ __forceinline__ void mx_bits_first_step(const AbMxFused &fx, int d, int n, uint32_t q, const uint32_t *__restrict__ tbuf,
                                                   const uint32_t *__restrict__ lut, uint32_t *sm) {
    constexpr int H2 = (P + 1) / 2, SUB = 1 << H2;
    const uint32_t qs = fx.dig[d].q_src, Hlo = (uint32_t)fx.dig[d].H;
    const int sh = 31 - __clz((int)fx.dig[d].pw);
    const int items = n / P;
    for (int v = threadIdx.x; v < items; v += blockDim.x) {
        const uint32_t *src = tbuf + v * P;
        uint32_t w[P];
        if (P % 4 == 0) {
            mx_static_for<(P % 4 == 0 ? P / 4 : 0)>([&](auto K) {
                constexpr int k = decltype(K)::value;
                const uint4 t = __ldcg(reinterpret_cast<const uint4 *>(src) + k);
                w[(4 * k) % P] = t.x; w[(4 * k + 1) % P] = t.y; w[(4 * k + 2) % P] = t.z; w[(4 * k + 3) % P] = t.w;
            });
        } else {
            mx_static_for<(P % 2 == 0 ? P / 2 : 0)>([&](auto K) {
                constexpr int k = decltype(K)::value;
                const uint2 t = __ldcg(reinterpret_cast<const uint2 *>(src) + k);
                w[(2 * k) % P] = t.x; w[(2 * k + 1) % P] = t.y;
            });
        }
        uint32_t m0 = 0, m1 = 0;       // bit c set <=> digit of column c is -1 <=> bit d of t is clear
        mx_static_for<P>([&](auto C) {
            constexpr int c = decltype(C)::value;
            const uint32_t lo = (uint32_t)ab_center(w[c], qs) + Hlo;
            const uint32_t neg = ((lo >> sh) & 1u) ^ 1u;
            if (c < H2) m0 |= neg << (c % H2);
            else m1 |= neg << ((c - H2) % H2);
        });
        uint32_t o[P];
        mx_static_for<P>([&](auto R) {
            constexpr int r = decltype(R)::value;
            o[r] = ab_addmod(__ldg(lut + r * SUB + m0), __ldg(lut + (P + r) * SUB + m1), q);
        });
        uint32_t *e = sm + v * P;
        if (P % 4 == 0) {
            mx_static_for<(P % 4 == 0 ? P / 4 : 0)>([&](auto K) {
                constexpr int k = decltype(K)::value;
                reinterpret_cast<uint4 *>(e)[k] = make_uint4(o[(4 * k) % P], o[(4 * k + 1) % P], o[(4 * k + 2) % P], o[(4 * k + 3) % P]);
            });
        } else {
            mx_static_for<(P % 2 == 0 ? P / 2 : 0)>([&](auto K) {
                constexpr int k = decltype(K)::value;
                reinterpret_cast<uint2 *>(e)[k] = make_uint2(o[(2 * k) % P], o[(2 * k + 1) % P]);
            });
        }
    }
}
__device__ __forceinline__ bool mx_bits_dispatch(int P, const AbMxFused &fx, int d, int n, uint32_t q, const uint32_t *tbuf, const uint32_t *lut, uint32_t *sm) {
    switch (P) {
        case 2: mx_bits_first_step<2>(fx, d, n, q, tbuf, lut, sm); return true;
        case 4: mx_bits_first_step<4>(fx, d, n, q, tbuf, lut, sm); return true;
        case 6: mx_bits_first_step<6>(fx, d, n, q, tbuf, lut, sm); return true;
        case 10: mx_bits_first_step<10>(fx, d, n, q, tbuf, lut, sm); return true;
        case 12: mx_bits_first_step<12>(fx, d, n, q, tbuf, lut, sm); return true;
        default: return false;
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

// CTA size: every pass hands ceil(items / T) work items to each thread and ends in a barrier, so a T that does not
// divide the item counts leaves threads idle in the last round (n = 9216, 12-point steps, two columns per item:
// 384 items -> 1.5 rounds at T = 256).  Model: time of one transform ~ sum over passes of rounds * work per item,
// throughput ~ resident CTAs / time; resident CTAs from shared memory and the 64-register budget (1024 threads).
// Measured on B200 (tools/sweep_crt.py, AB_MX_T): within 3 % of the best of {128, 160, 192, 224, 256} on every example ring.
static inline int pick_threads(const std::vector<AbMxStep> &steps, int n, size_t smem, int max_threads = 1024) {
    int best_t = 256;
    double best = 0;
    for (int T : {128, 160, 192, 256}) {
        const int ctas = (int)std::min<size_t>((size_t)(max_threads / T), (size_t)(227 * 1024) / (smem + 1024));
        if (ctas < 1) continue;
        double time = 6.0 * ((n / 2 + T - 1) / T);                 // load + store phases
        for (const AbMxStep &s : steps) {
            const int items = s.nwork / s.vec, rounds = (items + T - 1) / T;
            const double w = s.kind == 0 ? (double)s.span * s.vec * (s.span + 6) : s.kind == 2 ? (double)s.span * s.vec * (s.span / 2 + 12)
                                         : (double)s.span * s.vec * (3.5 * s.g + 2);
            time += rounds * w;
        }
        const double tput = ctas / time;
        if (tput > best * 1.0001) { best = tput; best_t = T; }
    }
    return best_t;
}

static inline int ab_mx_interleave() {
    // A/B measurements only.  Measured on B200: HomomRLWR mul_ level 1 +3 %, the CRT sweep unchanged except m = 27300 (five limbs,
    // n = 5760) -12 %: left off
    static const int v = (getenv("AB_MX_INTERLEAVE") && atoi(getenv("AB_MX_INTERLEAVE"))) ? 1 : 0;
    return v;
}

static inline uint32_t neg_inv32(uint32_t q) {   // -q^-1 mod 2^32 (Newton)
    uint32_t x = q;                        // q odd: x = q is an inverse mod 8
    for (int i = 0; i < 5; ++i) x *= 2u - q * x;
    return 0u - x;
}

static inline void fill_mx_args(const ab_plan *p, MxArgs &args, int32_t limb0, int32_t nlimbs, bool inverse, int &max_words) {
    const auto &steps = inverse ? p->mx_inv : p->mx_fwd;
    args.nsteps = (int32_t)steps.size();
    for (size_t i = 0; i < steps.size(); ++i) args.steps[i] = steps[i];
    args.n = (int32_t)p->n;
    args.tab_words[0] = inverse ? p->mx_words_inv[0] : p->mx_words_fwd[0];
    args.tab_words[1] = inverse ? p->mx_words_inv[1] : p->mx_words_fwd[1];
    args.nlimbs = nlimbs;
    args.interleave = ab_mx_interleave();
    max_words = 0;
    for (int l = 0; l < nlimbs; ++l) {
        const AbLimbDev &ld = p->limb[limb0 + l];
        args.limb[l].q = ld.q;
        args.limb[l].qinv_neg = neg_inv32(ld.q);
        args.limb[l].tab = inverse ? p->mx_tab_inv[limb0 + l] : p->mx_tab_fwd[limb0 + l];
        args.limb[l].wide = p->mx_ok[limb0 + l] == 2;
        {   // k (q-1)^2 + (2^32 - 1) q < 2^64 (the sum plus the Montgomery addend) and k q < 2^33 (reduced value below 3q: two conditional
            // subtractions): 8 terms for q < 2^30, 4 for the example moduli around 1.5e9, 2 just below 2^31
            const unsigned __int128 q = ld.q, one = 1;
            int k = 1;
            while (k < 8 && (unsigned __int128)(k + 1) * (q - 1) * (q - 1) + ((one << 32) - 1) * q < (one << 64) && (unsigned __int128)(k + 1) * q < (one << 33)) ++k;
            args.limb[l].lazy = k;
        }
        max_words = std::max(max_words, args.tab_words[args.limb[l].wide]);
    }
}

