// ab_mulks_pow2.cu -- the whole PT2CT `mul_` pipeline  modSwitch . keySwitchQuad hint . modSwitch . (*)
// (Crypto/Alchemy/Interpreter/PT2CT.hs:172-177; the operations of Eval.hs:65-67,130,133) as ONE kernel for
// power-of-two rings that fit shared memory -- the ring of examples/Arithmetic.hs (m' = 512, n = 256; BASELINE
// config 1).  One CTA owns one pair of input ciphertexts: it reads 4 L_in n residues, writes 2 L_out n, and every
// intermediate (the degree-2 product, the powerful-basis copy of c2, the gadget digits, the rescale corrections)
// lives in shared memory as uint32.  That is exactly the algorithmic traffic of SURVEY.md section 8(d):
// 32 n L_in + 16 n L_out bytes per op, hints amortised over the batch (they stay in L2).
//
// Steps (same order and arithmetic as the separate kernels of ab_she.cpp, so results are bit-identical):
//   A  (c0,c1,c2) = (a0 b0, a0 b1 + a1 b0, a1 b1)              (g = 1: a power of two has no odd prime), times the caller's
//      per-limb scalar (Lol toMSD: p^-1 mod q_l) when one is given
//   B  modSwitch up L_in -> L_h: multiply by the new moduli, new limbs are zero
//   C  TrivGad key switch: c2 -> powerful basis per limb (crtInv); digit d = centred limb d, reduced into every
//      other modulus and transformed (digit d in limb d is c2's CRT form itself); (c0,c1) += sum_d digit_d * hint_d
//   D  modSwitch down L_h -> L_out, one limb at a time: crtInv of the dropped limb, centred lift into each kept
//      modulus, CRT, out = (in - x) q_d^-1
// Transforms are plain radix-2 networks over shared memory with the twiddle order of ab_ntt_pow2.cu (stage s pairs
// (a, a + 2^s), twiddle tw[2^s + a mod 2^s]); several independent transforms share each pass and its barrier.
#include <algorithm>
#include <cstdlib>

#include "ab_internal.h"
#include "ab_modarith.cuh"

struct MkArgs {
    struct {
        uint32_t q, ninv, ninv_s, up, up_s, pad;
        uint64_t mu;
        const uint2 *fwd, *inv;
    } limb[AB_MAX_LIMBS];
    uint32_t dinv[AB_MAX_LIMBS][AB_MAX_LIMBS], dinv_s[AB_MAX_LIMBS][AB_MAX_LIMBS];   // [d][l] = q_d^-1 mod q_l
    int32_t logn, lin, lh, lout;
    int32_t scaled, pad;       // the product is multiplied by limb[l].up: appended moduli (L_h > L_in) and / or the caller's per-limb scalar (toMSD)
    int64_t batch;
};

// `count` transforms of n points; transform t lives at buf + t * n and uses modulus limb_of(t); skip(t) leaves it alone.
// Two stages per pass (four residues per work item: half the shared-memory traffic, index arithmetic and barriers of
// a stage-per-pass network); an odd stage count ends (forward) or starts (inverse) with one single stage.
__device__ __forceinline__ void mk_bf_fwd(uint32_t &X, uint32_t &Y, uint2 w, uint32_t q) {
    const uint32_t t = ab_shoup(Y, w.x, w.y, q), x = X;
    X = ab_addmod(x, t, q);
    Y = ab_submod(x, t, q);
}
__device__ __forceinline__ void mk_bf_inv(uint32_t &X, uint32_t &Y, uint2 w, uint32_t q) {
    const uint32_t s = ab_addmod(X, Y, q), d = ab_submod(X, Y, q);
    X = s;
    Y = ab_shoup(d, w.x, w.y, q);
}
template <bool INV, class LimbOf, class Skip>
__device__ __forceinline__ void ntt_batch(const MkArgs &a, uint32_t *buf, int count, LimbOf limb_of, Skip skip) {
    const int logn = a.logn;
    auto single = [&](int s) {
        const int h = 1 << s, half = 1 << (logn - 1), total = count << (logn - 1);
        for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
            const int t = idx >> (logn - 1), b = idx & (half - 1);
            if (skip(t)) continue;
            const int l = limb_of(t);
            const uint32_t q = a.limb[l].q;
            const int lo = b & (h - 1);
            uint32_t *x = buf + (t << logn) + (((b >> s) << (s + 1)) | lo);
            const uint2 w = __ldg((INV ? a.limb[l].inv : a.limb[l].fwd) + h + lo);
            uint32_t X = x[0], Y = x[h];
            if (!INV) mk_bf_fwd(X, Y, w, q);
            else {
                mk_bf_inv(X, Y, w, q);
                if (s == 0) X = ab_shoup(X, a.limb[l].ninv, a.limb[l].ninv_s, q);      // tw[1] already carries n^-1
            }
            x[0] = X; x[h] = Y;
        }
        __syncthreads();
    };
    auto pair = [&](int s) {            // stages s and s + 1
        const int h = 1 << s, quarter = 1 << (logn - 2), total = count << (logn - 2);
        for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
            const int t = idx >> (logn - 2), g = idx & (quarter - 1);
            if (skip(t)) continue;
            const int l = limb_of(t);
            const uint32_t q = a.limb[l].q;
            const int lo = g & (h - 1);
            uint32_t *x = buf + (t << logn) + (((g >> s) << (s + 2)) | lo);
            const uint2 *tw = INV ? a.limb[l].inv : a.limb[l].fwd;
            const uint2 w1 = __ldg(tw + h + lo), w2a = __ldg(tw + 2 * h + lo), w2b = __ldg(tw + 3 * h + lo);
            uint32_t x0 = x[0], x1 = x[h], x2 = x[2 * h], x3 = x[3 * h];
            if (!INV) {
                mk_bf_fwd(x0, x1, w1, q); mk_bf_fwd(x2, x3, w1, q);
                mk_bf_fwd(x0, x2, w2a, q); mk_bf_fwd(x1, x3, w2b, q);
            } else {
                mk_bf_inv(x0, x2, w2a, q); mk_bf_inv(x1, x3, w2b, q);
                mk_bf_inv(x0, x1, w1, q); mk_bf_inv(x2, x3, w1, q);
                if (s == 0) {
                    x0 = ab_shoup(x0, a.limb[l].ninv, a.limb[l].ninv_s, q);
                    x2 = ab_shoup(x2, a.limb[l].ninv, a.limb[l].ninv_s, q);
                }
            }
            x[0] = x0; x[h] = x1; x[2 * h] = x2; x[3 * h] = x3;
        }
        __syncthreads();
    };
    if (!INV) {
        int s = 0;
        for (; s + 1 < logn; s += 2) pair(s);
        if (s < logn) single(s);
    } else {
        int s = logn - 1;
        if ((logn & 1) != 0) { single(s); --s; }
        for (; s >= 1; s -= 2) pair(s - 1);
    }
}

__device__ __forceinline__ uint32_t mk_reduce_centered(uint32_t v, uint32_t qs, uint32_t q) {
    const int32_t c = ab_center(v, qs);
    uint32_t m = c < 0 ? (uint32_t)(-c) : (uint32_t)c;
    if (m >= q) { m -= q; if (m >= q) m %= q; }
    return (c < 0 && m) ? q - m : m;
}

__global__ void __launch_bounds__(256) k_mulks_pow2(const __grid_constant__ MkArgs a, int64_t *__restrict__ out2, const int64_t *__restrict__ a2,
                                                    const int64_t *__restrict__ b2, const int64_t *__restrict__ hint) {
    extern __shared__ __align__(16) uint32_t smem[];
    const int n = 1 << a.logn, lin = a.lin, lh = a.lh, lout = a.lout;
    uint32_t *c = smem;                         // [3][lh][n]
    uint32_t *pw = c + 3 * lh * n;              // [lin][n]   c2 in the powerful basis, per limb
    uint32_t *sc = pw + (lin > 2 ? lin : 2) * n; // scratch [max(lh, 2 (lh - 1))][n]  (step D parks two polynomials in pw)
    const int64_t b = blockIdx.x;
    const int T = blockDim.x, tid = threadIdx.x;

    // ---- A + B: product in the secret, scaled up to the hint modulus
    const int64_t *ap = a2 + b * 2 * lin * n, *bp = b2 + b * 2 * lin * n;
    for (int i = tid; i < lin * n; i += T) {
        const int l = i >> a.logn;
        const uint32_t q = a.limb[l].q;
        const uint64_t mu = a.limb[l].mu;
        const uint32_t a0 = (uint32_t)__ldcs(ap + i), a1 = (uint32_t)__ldcs(ap + lin * n + i);
        const uint32_t b0 = (uint32_t)__ldcs(bp + i), b1 = (uint32_t)__ldcs(bp + lin * n + i);
        uint32_t c0 = ab_mulmod(a0, b0, q, mu);
        uint32_t c1 = ab_addmod(ab_mulmod(a0, b1, q, mu), ab_mulmod(a1, b0, q, mu), q);
        uint32_t c2 = ab_mulmod(a1, b1, q, mu);
        if (a.scaled) {
            c0 = ab_shoup(c0, a.limb[l].up, a.limb[l].up_s, q);
            c1 = ab_shoup(c1, a.limb[l].up, a.limb[l].up_s, q);
            c2 = ab_shoup(c2, a.limb[l].up, a.limb[l].up_s, q);
        }
        c[i] = c0; c[lh * n + i] = c1; c[2 * lh * n + i] = c2; pw[i] = c2;
    }
    for (int i = lin * n + tid; i < lh * n; i += T) { c[i] = 0; c[lh * n + i] = 0; }     // new limbs of c0, c1 (c2's are never read)
    __syncthreads();

    // ---- C: key switch.  c2 -> powerful basis in every limb that carries data
    ntt_batch<true>(a, pw, lin, [](int t) { return t; }, [](int) { return false; });
    for (int d = 0; d < lin; ++d) {
        const uint32_t qs = a.limb[d].q;
        for (int i = tid; i < lh * n; i += T) {
            const int l = i >> a.logn, j = i & (n - 1);
            sc[i] = (l == d) ? c[2 * lh * n + d * n + j] : mk_reduce_centered(pw[d * n + j], qs, a.limb[l].q);
        }
        __syncthreads();
        ntt_batch<false>(a, sc, lh, [](int t) { return t; }, [d](int t) { return t == d; });
        const int64_t *h0 = hint + (int64_t)(2 * d) * lh * n, *h1 = h0 + (int64_t)lh * n;
        for (int i = tid; i < lh * n; i += T) {
            const int l = i >> a.logn;
            const uint32_t q = a.limb[l].q, x = sc[i];
            const uint64_t mu = a.limb[l].mu;
            c[i] = ab_addmod(c[i], ab_mulmod(x, (uint32_t)__ldg(h0 + i), q, mu), q);
            c[lh * n + i] = ab_addmod(c[lh * n + i], ab_mulmod(x, (uint32_t)__ldg(h1 + i), q, mu), q);
        }
        __syncthreads();
    }

    // ---- D: rescale down, one limb at a time.  x_i = crtInv(c_i[d]) (in place), t[i][l] = CRT_l(lift x_i)
    for (int d = lh - 1; d >= lout; --d) {
        // the two dropped-limb polynomials sit lh*n apart: transform them as a batch of two with stride lh*n -> copy to pw first
        uint32_t *xd = pw;      // [2][n] (pw is dead after step C; lin >= 1 and we need 2 n <= (lin + scratch) n: see host check)
        for (int i = tid; i < 2 * n; i += T) xd[i] = c[(i >> a.logn) * lh * n + d * n + (i & (n - 1))];
        __syncthreads();
        ntt_batch<true>(a, xd, 2, [d](int) { return d; }, [](int) { return false; });
        const uint32_t qs = a.limb[d].q;
        for (int i = tid; i < 2 * d * n; i += T) {
            const int t = i >> a.logn, j = i & (n - 1), poly = t / d, l = t - poly * d;
            sc[i] = mk_reduce_centered(xd[poly * n + j], qs, a.limb[l].q);
        }
        __syncthreads();
        ntt_batch<false>(a, sc, 2 * d, [d](int t) { return t % d; }, [](int) { return false; });
        for (int i = tid; i < 2 * d * n; i += T) {
            const int t = i >> a.logn, j = i & (n - 1), poly = t / d, l = t - poly * d;
            const uint32_t q = a.limb[l].q;
            uint32_t *cp = c + poly * lh * n + l * n + j;
            *cp = ab_shoup(ab_submod(*cp, sc[i], q), a.dinv[d][l], a.dinv_s[d][l], q);
        }
        __syncthreads();
    }

    int64_t *op = out2 + b * 2 * lout * n;
    for (int i = tid; i < 2 * lout * n; i += T) {
        const int poly = i / (lout * n), r = i - poly * lout * n;
        __stcs(op + i, (int64_t)c[poly * lh * n + r]);
    }
}

static size_t mulks_smem(const ab_plan *p, int32_t lin, int32_t lh) {
    const int64_t scratch = std::max<int64_t>(lh, 2 * (int64_t)(lh - 1));
    const int64_t pwn = std::max<int64_t>(lin, 2);            // step D parks two polynomials in pw
    return (size_t)((3 * (int64_t)lh + pwn + scratch) * p->n * (int64_t)sizeof(uint32_t));
}

bool ab_mulks_pow2_supported(const ab_plan *p, int32_t lin, int32_t lh, int32_t lout, int32_t base) {
    static const bool off = getenv("AB_NO_MULKS") && atoi(getenv("AB_NO_MULKS")) != 0;      // A/B measurements only
    if (off || !p->pow2 || !p->has_crt || p->n < 2 || base != 0) return false;
    if (lin < 1 || lh < lin || lout < 1 || lout > lh || lh > p->L) return false;
    for (int l = 0; l < lh; ++l) if (!p->limb[l].tw_fwd_flat) return false;
    return mulks_smem(p, lin, lh) <= 100 * 1024;       // at least two CTAs per SM
}

void ab_launch_mulks_pow2(const ab_plan *p, int64_t *out2, const int64_t *a2, const int64_t *b2, const int64_t *hint, int64_t batch,
                          int32_t lin, int32_t lh, int32_t lout, cudaStream_t st, const int64_t *prescale) {
    AB_REQUIRE(ab_mulks_pow2_supported(p, lin, lh, lout, 0), AB_ERR_UNSUPPORTED, "fused mul + key switch: not supported for this plan");
    if (batch == 0) return;
    MkArgs a{};
    a.logn = p->logn; a.lin = lin; a.lh = lh; a.lout = lout; a.batch = batch; a.scaled = (lh > lin || prescale) ? 1 : 0;
    for (int l = 0; l < lh; ++l) {
        const AbLimbDev &ld = p->limb[l];
        const uint64_t q = (uint64_t)p->q[l];
        a.limb[l].q = ld.q; a.limb[l].ninv = ld.ninv; a.limb[l].ninv_s = ld.ninv_s; a.limb[l].mu = ld.mu;
        a.limb[l].fwd = reinterpret_cast<const uint2 *>(ld.tw_fwd_flat);
        a.limb[l].inv = reinterpret_cast<const uint2 *>(ld.tw_inv_flat);
        uint64_t up = 1;
        for (int k = lin; k < lh; ++k) up = abh::mulmod(up, (uint64_t)p->q[k] % q, q);
        if (prescale && l < lin) up = abh::mulmod(up, (uint64_t)(((prescale[l] % (int64_t)q) + (int64_t)q) % (int64_t)q), q);
        a.limb[l].up = (uint32_t)up; a.limb[l].up_s = abh::shoup(up, q);
        for (int d = 0; d < lh; ++d)
            if (d != l) {
                const uint64_t inv = abh::invmod((uint64_t)p->q[d] % q, q);
                a.dinv[d][l] = (uint32_t)inv; a.dinv_s[d][l] = abh::shoup(inv, q);
            }
    }
    const size_t smem = mulks_smem(p, lin, lh);
    static thread_local bool ready = false;
    if (!ready) {
        AB_CHECK_CUDA(cudaFuncSetAttribute(k_mulks_pow2, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        ready = true;
    }
    const int T = (int)std::max<int64_t>(64, std::min<int64_t>(256, p->n / 2));
    k_mulks_pow2<<<(unsigned)batch, T, smem, st>>>(a, out2, a2, b2, hint);
    ab_count_launch();
    ab_check_launch("mulks_pow2");
}
