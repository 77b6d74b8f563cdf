// ab_mulks_n256.cu -- the PT2CT `mul_` pipeline  modSwitch . keySwitchQuad hint . modSwitch . (*)  of examples/Arithmetic.hs
// (Crypto/Alchemy/Interpreter/PT2CT.hs:172-177 on the ring F512: m' = 512, n = 256, two limbs in, two-limb TrivGad hint, one limb
// out -- SURVEY.md B.3) as ONE kernel with ONE WARP per pair of input ciphertexts and every transform in registers.
//
// ab_mulks_pow2.cu runs the same pipeline for any power-of-two ring as radix-2 passes over shared memory (four residues per work
// item, a CTA barrier per pass): at n = 256 that kernel is instruction-bound -- 15 k warp instructions per op, half of them in the
// 32 passes of its eight transforms, whose 4-byte accesses at strides 1 and 4 also cost 2.75 x the ideal shared-memory wavefronts.
// Here a transform is 16 lanes x 16 residues with the register windows of ab_ntt_pow2.cu (two windows of four stages, one padded
// shared-memory transpose between them), the two transforms every phase of this limb schedule has run on the two half-warps, and
// only __syncwarp separates the phases.  General products are Montgomery (one 64-bit multiply-add chain and one reduction per
// sum of two products; the factor 2^32 rides in with the caller's toMSD scalar), (c0, c1, c2) wait in a per-lane shared-memory stash
// while the transforms run, so that the kernel fits 80 registers and 24 warps per SM.  Measured (tools/bench_she.py, 32768 ops):
// 57.6 M ops/s on ab_mulks_pow2.cu, 155 M here; 4.0 k warp instructions per op; bound by the integer-multiply pipe.
//
//   lanes 0-15 = limb 0, lanes 16-31 = limb 1 (until the rescale); "strided" = lane t holds positions t + 16 i, "contiguous" = 16 t + i
//   A  (c0, c1, c2) = (a0 b0, a0 b1 + a1 b0, a1 b1) times the caller's toMSD scalar        strided, straight from HBM
//   C  c2 -> powerful basis (inverse windows 1, 0); the halves swap them through shared memory; digit = centred lift of the OTHER
//      limb reduced into this one -> forward windows 0, 1; (c0, c1) += digit * hint_other + c2 * hint_own                strided
//   D  drop limb 1: its c0 / c1 go to the two halves, inverse transform, centred lift into q0, forward transform in limb 0,
//      out = (c - x) q1^-1                                                                  strided, straight to HBM
// Same arithmetic as the launch sequence and as ab_mulks_pow2.cu: results are bit-identical (tests/test_gpu_she.py).
#include <algorithm>
#include <cstdlib>

#include "ab_ntt_pow2_dev.cuh"

namespace {

#ifndef AB_M256_OPS
#define AB_M256_OPS 4
#endif
#ifndef AB_M256_MINB
#define AB_M256_MINB 6
#endif
#ifndef AB_M256_LDB
#define AB_M256_LDB 2
#endif
constexpr int LOGN = 8, LOGE = 4, N = 1 << LOGN, E = 1 << LOGE, PADN = N + (N >> 4), OPS = AB_M256_OPS;      // OPS warps (ops) per CTA
constexpr int LDB = AB_M256_LDB;                                                                   // residue pairs per lane whose four 16-byte loads are issued together in phase A
constexpr int MINB = AB_M256_MINB;                                                                 // CTAs per SM the register budget is cut for

struct M256Limb {
    uint32_t qinv;                    // q^-1 mod 2^32 (Montgomery)
    uint32_t qrec;                    // floor(2^32 / q): one-multiplication reduction of a lifted digit
    uint32_t r, r_s;                  // R = 2^32 mod q, Shoup pair: plain -> Montgomery form
    uint32_t upr, upr_s;              // (toMSD scalar) * R: one factor of every product carries it, the Montgomery reduction then lands on the scaled product
};
struct M256Args {
    Pow2Limb fwd[2], inv[2];          // window-packed twiddles of the two limbs, both directions
    M256Limb k[2];
    uint32_t d1inv, d1inv_s;          // q1^-1 mod q0
    int64_t batch, pf_dist;
};

#define M256_FW(W, lp, low) fwd_window<LOGE, W, ab_pow2_s(LOGN, LOGE, W), ab_pow2_lb0(LOGN, LOGE, W), true>(x, lp, lp.tw, ab_pow2_base(LOGN, LOGE, W), low, lp.q)
#define M256_IW(W, lp, low) inv_window<LOGE, W, ab_pow2_s(LOGN, LOGE, W), ab_pow2_lb0(LOGN, LOGE, W), true>(x, lp, lp.tw, ab_pow2_base(LOGN, LOGE, W), low, lp.q)

// padding of a half-warp's 256-word buffer: one word per 16, and the two halves 16 banks apart -- the strided pattern (t + 16 i)
// and the contiguous one (16 t + i) of one half each cover 16 distinct banks, and the other half's accesses fall on the other 16
__device__ __forceinline__ int ph(int a) { return a + (a >> 4); }

// contiguous (fully reduced or below 4q) -> strided, below 4q.  T: this half-warp's transpose buffer
__device__ __forceinline__ void ntt_fwd(uint32_t (&x)[E], const Pow2Limb &lp, uint32_t *T, int tl) {
    M256_FW(0, lp, 0);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < E; ++i) T[ph(16 * tl + i)] = x[i];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < E; ++i) x[i] = T[ph(tl + 16 * i)];
    M256_FW(1, lp, tl);
}
// strided (below 2q) -> contiguous, below 2q, scaled by n^-1
__device__ __forceinline__ void ntt_inv(uint32_t (&x)[E], const Pow2Limb &lp, uint32_t *T, int tl) {
    M256_IW(1, lp, tl);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < E; ++i) T[ph(tl + 16 * i)] = x[i];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < E; ++i) x[i] = T[ph(16 * tl + i)];
    M256_IW(0, lp, 0);
}
// centred lift of v mod qs (below 2^30), reduced into q: |lift| < 2^29, and floor(|lift| * qrec / 2^32) misses the quotient by at most one
__device__ __forceinline__ uint32_t lift_into(uint32_t v, uint32_t qs, uint32_t q, uint32_t qrec) {
    const int32_t c = ab_center(v, qs);
    const uint32_t m = c < 0 ? (uint32_t)(-c) : (uint32_t)c;
    uint32_t r = m;
    if (qs > 2 * q) r = ab_csub(m - __umulhi(m, qrec) * q, q);      // otherwise |lift| <= qs / 2 <= q already (uniform per half-warp)
    return (c < 0 && r) ? q - r : r;
}
// Montgomery reduction: T 2^-32 mod q in [0, q) for T < q 2^32
__device__ __forceinline__ uint32_t redc(uint64_t T, uint32_t q, uint32_t qinv) {
    const uint32_t t = (uint32_t)(T >> 32) - __umulhi((uint32_t)T * qinv, q);
    return min(t, t + q);
}

// The straight-line code of four 256-point transforms plus sixteen-fold unrolled element loops is 64 KB, twice the instruction cache
// of an SM, and with 24 warps at different places instruction fetch became the first stall reason (ncu: no_instruction 3.7 per issued
// instruction).  So: the element loops run rolled over shared memory (only the transforms need compile-time register indices), and
// phases C and D, which are both  inverse transform -> centred lift into another modulus -> forward transform,  share one loop body.
__global__ void __launch_bounds__(32 * OPS, MINB) k_mulks_n256(const __grid_constant__ M256Args a, int64_t *__restrict__ out2, const int64_t *__restrict__ a2,
                                                               const int64_t *__restrict__ b2, const int64_t *__restrict__ hint) {
    // per warp: a transpose / exchange buffer per half-warp, and the stash of (c0, c1, c2) at [polynomial][i][lane]: a lane only
    // ever touches its own column, except where the two halves hand each other a polynomial in phase D
    __shared__ uint32_t s_t[OPS][2][PADN];
    __shared__ uint32_t s_c[OPS][3][E][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, half = lane >> 4, tl = lane & 15;
    uint32_t *T = s_t[warp][half], *To = s_t[warp][half ^ 1];
    uint32_t (*C)[E][32] = s_c[warp];
    const int l = half;                                         // this half-warp's limb in phases A and C
    const uint32_t q = a.fwd[l].q;
    const M256Limb kl = a.k[l];

    const int64_t b = (int64_t)blockIdx.x * OPS + warp;
    if (b >= a.batch) return;                                   // whole warps only: the phases below synchronise per warp
    // pull the inputs of an op that starts an eighth of a wave later into L2 (two 8 KB bulk prefetches per warp): +3.7 % (any distance
    // from 0.03 to 0.25 waves measured the same, one wave nothing, two waves -9 %: the lines fall out of L2 again)
    if (tl == 0 && b + a.pf_dist < a.batch) {
        const int64_t *nx = (half ? b2 : a2) + (b + a.pf_dist) * (2 * 2 * N);
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nx), "r"(2 * 2 * N * 8) : "memory");
    }

    // ---- A: product in the secret, straight from HBM; Montgomery products, 2^32 rides in with the toMSD scalar
    {
        // 16-byte loads: a lane takes the residue PAIRS at 2 t + 32 k, so a half-warp request is 256 contiguous bytes (with 8-byte loads
        // at t + 16 i it was 128, and the loads, products and stores alone ran at 0.70 of the HBM peak; residues are below 2^31, only
        // the low words are used).  The products are pointwise, so any assignment of positions to lanes will do here: they go to the
        // stash slot of the lane that owns the position in the strided layout of the transforms, (i, lane) = (j / 16, j % 16).
        const int4 *ap = reinterpret_cast<const int4 *>(a2 + (b * 2 * 2 + l) * N) + tl;      // [b][poly][limb][j]; poly 1 is 2 N residues further
        const int4 *bp = reinterpret_cast<const int4 *>(b2 + (b * 2 * 2 + l) * N) + tl;
        uint32_t *cs = &C[0][tl >> 3][16 * half + ((2 * tl) & 15)];                          // + 2 k rows; the three polynomials are E rows apart
#pragma unroll 1
        for (int k0 = 0; k0 < E / 2; k0 += LDB) {
            int4 a0[LDB], a1[LDB], b0[LDB], b1[LDB];
#pragma unroll
            for (int k = 0; k < LDB; ++k) {
                const int o = 16 * (k0 + k);
                a0[k] = __ldcs(ap + o); a1[k] = __ldcs(ap + N + o);
                b0[k] = __ldcs(bp + o); b1[k] = __ldcs(bp + N + o);
            }
#pragma unroll
            for (int k = 0; k < LDB; ++k) {
                uint32_t r[3][2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const uint32_t x0 = e ? a0[k].z : a0[k].x, x1 = e ? a1[k].z : a1[k].x, y0 = e ? b0[k].z : b0[k].x, y1 = e ? b1[k].z : b1[k].x;
                    const uint32_t s0 = ab_shoup(x0, kl.upr, kl.upr_s, q), s1 = ab_shoup(x1, kl.upr, kl.upr_s, q);      // a * scalar * 2^32
                    r[0][e] = redc((uint64_t)s0 * y0, q, kl.qinv);
                    r[1][e] = redc((uint64_t)s0 * y1 + (uint64_t)s1 * y0, q, kl.qinv);                                 // 2 q^2 < q 2^32
                    r[2][e] = redc((uint64_t)s1 * y1, q, kl.qinv);
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) *reinterpret_cast<uint2 *>(cs + (c * E + 2 * (k0 + k)) * 32) = make_uint2(r[c][0], r[c][1]);
            }
        }
        __syncwarp();                                           // the stash rows were written by other lanes of the half-warp
    }

    uint32_t x[E];
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        // pass 0 = C: c2 to the powerful basis in both limbs; each half takes the OTHER's coefficients as its gadget digit (TrivGad)
        // pass 1 = D: drop limb 1 -- half 0 takes c0 of limb 1 from half 1's columns, half 1 keeps its c1; both lift into limb 0
        const int li = pass ? 1 : l, lf = pass ? 0 : l;
        const uint32_t qi = a.inv[li].q, qs = a.inv[pass ? 1 : (l ^ 1)].q, qf = a.fwd[lf].q;
        if (pass == 0) {
#pragma unroll
            for (int i = 0; i < E; ++i) x[i] = C[2][i][lane];
        } else {
            __syncwarp();
#pragma unroll
            for (int i = 0; i < E; ++i) x[i] = C[half][i][16 + tl];
        }
        ntt_inv(x, a.inv[li], T, tl);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < E; ++i) T[ph(16 * tl + i)] = ab_csub(x[i], qi);
        __syncwarp();
        const uint32_t *src = pass ? T : To;
        const uint32_t qrec = a.k[lf].qrec;
#pragma unroll
        for (int i = 0; i < E; ++i) x[i] = lift_into(src[ph(16 * tl + i)], qs, qf, qrec);
        ntt_fwd(x, a.fwd[lf], T, tl);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < E; ++i) T[ph(tl + 16 * i)] = x[i];                  // read back by this lane only, in the rolled loops below
        if (pass == 0) {
            // hint[d][component][limb][j]: the digit of the other limb (d = l ^ 1) and this limb's own (d = l: its transform is c2 itself)
            const int64_t *ho = hint + ((int64_t)((l ^ 1) * 2) * 2 + l) * N + tl, *ht = hint + ((int64_t)(l * 2) * 2 + l) * N + tl;
#pragma unroll 4
            for (int i = 0; i < E; ++i) {
                const uint32_t d = ab_shoup(T[ph(tl + 16 * i)], kl.r, kl.r_s, q);                  // any 32-bit input; the digit in Montgomery form
                const uint32_t c2m = ab_shoup(C[2][i][lane], kl.r, kl.r_s, q);
                const uint32_t h0o = (uint32_t)__ldg(ho + 16 * i), h1o = (uint32_t)__ldg(ho + 2 * N + 16 * i);
                const uint32_t h0t = (uint32_t)__ldg(ht + 16 * i), h1t = (uint32_t)__ldg(ht + 2 * N + 16 * i);
                C[0][i][lane] = ab_addmod(C[0][i][lane], redc((uint64_t)d * h0o + (uint64_t)c2m * h0t, q, kl.qinv), q);
                C[1][i][lane] = ab_addmod(C[1][i][lane], redc((uint64_t)d * h1o + (uint64_t)c2m * h1t, q, kl.qinv), q);
            }
        }
    }

    const uint32_t q0 = a.fwd[0].q;
    int64_t *op = out2 + (b * 2 + half) * N + tl;               // [b][poly][1 limb][j]
#pragma unroll 4
    for (int i = 0; i < E; ++i) {
        const uint32_t cc = C[half][i][tl];                     // limb 0: polynomial `half` sits in half 0's columns
        const uint32_t xi = T[ph(tl + 16 * i)];
        const uint32_t corr = ab_csub(ab_csub(xi, 2 * q0), q0);
        __stcs(op + 16 * i, (int64_t)ab_shoup(ab_submod(cc, corr, q0), a.d1inv, a.d1inv_s, q0));
    }
}

}   // namespace

bool ab_mulks_n256_supported(const ab_plan *p, int32_t lin, int32_t lh, int32_t lout, int32_t base) {
    static const bool off = (getenv("AB_NO_MULKS256") && atoi(getenv("AB_NO_MULKS256")) != 0) ||
                            (getenv("AB_NO_MULKS") && atoi(getenv("AB_NO_MULKS")) != 0);      // A/B measurements only
    if (off || !p->pow2 || !p->has_crt || p->logn != LOGN || base != 0 || lin != 2 || lh != 2 || lout != 1 || p->L < 2) return false;
    for (int l = 0; l < 2; ++l)
        if (p->limb[l].q >= (1u << 30) || !p->limb[l].tw_fwd || !p->limb[l].tw_inv) return false;      // Harvey butterflies
    return ab_pow2_loge(LOGN) == LOGE;
}

void ab_launch_mulks_n256(const ab_plan *p, int64_t *out2, const int64_t *a2, const int64_t *b2, const int64_t *hint, int64_t batch,
                          cudaStream_t st, const int64_t *prescale) {
    AB_REQUIRE(ab_mulks_n256_supported(p, 2, 2, 1, 0), AB_ERR_UNSUPPORTED, "n = 256 mul + key switch kernel: not supported for this plan");
    if (batch == 0) return;
    M256Args a{};
    a.batch = batch;
    a.pf_dist = std::max<int64_t>(OPS, (int64_t)p->sm_count * MINB * OPS / 8);      // ops ahead of the one a warp works on: an eighth of a wave
    for (int l = 0; l < 2; ++l) {
        const AbLimbDev &ld = p->limb[l];
        const uint64_t q = (uint64_t)p->q[l];
        for (int dir = 0; dir < 2; ++dir) {
            Pow2Limb &pl = dir ? a.inv[l] : a.fwd[l];
            pl.q = ld.q; pl.ninv = ld.ninv; pl.ninv_s = ld.ninv_s;
            pl.tw = reinterpret_cast<const uint2 *>(dir ? ld.tw_inv : ld.tw_fwd);
            const uint2_ *t0 = dir ? ld.tw0_inv : ld.tw0_fwd;
            for (int k = 0; k < 32; ++k) pl.tw0[k] = make_uint2(t0[k].x, t0[k].y);
        }
        M256Limb &k = a.k[l];
        uint32_t qi = (uint32_t)q;                                // Newton: five steps double the correct bits from 3 to 96
        for (int it = 0; it < 5; ++it) qi *= 2u - (uint32_t)q * qi;
        k.qinv = qi;
        k.qrec = (uint32_t)((1ull << 32) / q);
        const uint64_t up = prescale ? (uint64_t)(((prescale[l] % (int64_t)q) + (int64_t)q) % (int64_t)q) : 1;
        const uint64_t R = (1ull << 32) % q, upr = up * R % q;
        k.r = (uint32_t)R; k.r_s = abh::shoup(R, q);
        k.upr = (uint32_t)upr; k.upr_s = abh::shoup(upr, q);
    }
    const uint64_t q0 = (uint64_t)p->q[0], inv = abh::invmod((uint64_t)p->q[1] % q0, q0);
    a.d1inv = (uint32_t)inv; a.d1inv_s = abh::shoup(inv, q0);
    k_mulks_n256<<<(unsigned)((batch + OPS - 1) / OPS), 32 * OPS, 0, st>>>(a, out2, a2, b2, hint);
    ab_count_launch();
    ab_check_launch("mulks_n256");
}
