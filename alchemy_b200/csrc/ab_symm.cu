// ab_symm.cu -- the ciphertext <-> plaintext boundary of Lol's SymmSHE on the device (SURVEY.md section 8 rows
// a9-a12 and f3): what ALCHEMY reaches through addPublic_/mulPublic_ (Crypto/Alchemy/Interpreter/Eval.hs:131-132),
// encrypt/decrypt (Crypto/Alchemy/Interpreter/PT2CT.hs:78-99) and errorRate_ (Eval.hs:155-160).
// Upstream does all of this in Haskell over Cyc values [RECALL]; here each piece is one streaming kernel over a
// batch of ciphertexts, planar layout, one coalesced read/write per residue:
//   * ab_ct_phase_dev          c(s) = c0 + c1 s + c2 s^2 ... in the CRT basis (Horner) -- `evaluate c sq` of errorTerm
//   * ab_rns_lift_reduce_dev   centred lift of every coefficient over the RNS product (Garner mixed radix, no big
//                              integers), reduced mod the plaintext modulus; max |lift| / Q per element (errorRate)
//   * ab_lift_plain_dev        reduce . lift of a plaintext coefficient vector into every ciphertext modulus
//   * ab_ct_add_public_dev / ab_ct_mul_public_dev    the pointwise cores of addPublic / mulPublic
#include <vector>

#include "ab_internal.h"
#include "ab_modarith.cuh"

static inline unsigned symm_grid(int64_t work, int sm_count) {
    int64_t g = (work + 255) / 256, cap = (int64_t)sm_count * 16;
    return (unsigned)std::max<int64_t>(1, std::min(g, cap));
}

// ------------------------------------------------------------------------------------ phase
struct PhArgs {
    struct { uint32_t q, pad; uint64_t mu; } limb[AB_MAX_LIMBS];
    int32_t nlimbs, npoly;
    int64_t n, batch;
};
__global__ void __launch_bounds__(256) k_ct_phase(const PhArgs a, int64_t *__restrict__ out, const int64_t *__restrict__ c, const int64_t *__restrict__ s) {
    const int64_t per = (int64_t)a.nlimbs * a.n, total = a.batch * per;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / per, r = i % per;
        const int l = (int)(r / a.n);
        const uint32_t q = a.limb[l].q, sv = (uint32_t)__ldg(s + r);
        const int64_t *cp = c + b * a.npoly * per + r;
        uint32_t acc = (uint32_t)__ldcs(cp + (int64_t)(a.npoly - 1) * per);
        for (int k = a.npoly - 2; k >= 0; --k)
            acc = ab_addmod(ab_mulmod(acc, sv, q, a.limb[l].mu), (uint32_t)__ldcs(cp + (int64_t)k * per), q);
        out[i] = (int64_t)acc;
    }
}
extern "C" int ab_ct_phase_dev(const ab_plan *p, int64_t *out, const int64_t *c, const int64_t *s, int64_t batch, int32_t npoly, int32_t nlimbs, void *stream) {
    return ab_guard([&] {
        AB_REQUIRE(p && ((out && c && s) || batch == 0), AB_ERR_ARG, "null argument");
        AB_REQUIRE(batch >= 0 && npoly >= 1 && nlimbs >= 1 && nlimbs <= p->L, AB_ERR_ARG, "bad batch, degree or limb count");
        if (batch == 0) return;
        PhArgs a{};
        a.nlimbs = nlimbs; a.npoly = npoly; a.n = p->n; a.batch = batch;
        for (int l = 0; l < nlimbs; ++l) { a.limb[l].q = p->limb[l].q; a.limb[l].mu = p->limb[l].mu; }
        k_ct_phase<<<symm_grid(batch * nlimbs * p->n, p->sm_count), 256, 0, (cudaStream_t)stream>>>(a, out, c, s);
        ab_count_launch();
        ab_check_launch("ct_phase");
    });
}

// ------------------------------------------------------------------------------------ public operands
struct PubArgs {
    struct { uint32_t q, pad; uint64_t mu; } limb[AB_MAX_LIMBS];
    int32_t nlimbs, npoly, mul, pad;
    int64_t n, batch;
};
__global__ void __launch_bounds__(256) k_ct_public(const PubArgs a, int64_t *__restrict__ c, const int64_t *__restrict__ pub) {
    const int64_t per = (int64_t)a.nlimbs * a.n;
    const int64_t polys = a.mul ? a.npoly : 1, total = a.batch * polys * per;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / (polys * per), r = i % (polys * per), rr = r % per;
        const int l = (int)(rr / a.n);
        const uint32_t q = a.limb[l].q, pv = (uint32_t)__ldg(pub + rr);
        int64_t *cp = c + b * a.npoly * per + r;
        const uint32_t v = (uint32_t)*cp;
        *cp = (int64_t)(a.mul ? ab_mulmod(v, pv, q, a.limb[l].mu) : ab_addmod(v, pv, q));
    }
}
static void ct_public(const ab_plan *p, int64_t *c, const int64_t *pub, int64_t batch, int32_t npoly, int32_t nlimbs, bool mul, void *stream) {
    AB_REQUIRE(p && ((c && pub) || batch == 0), AB_ERR_ARG, "null argument");
    AB_REQUIRE(batch >= 0 && npoly >= 1 && nlimbs >= 1 && nlimbs <= p->L, AB_ERR_ARG, "bad batch, degree or limb count");
    if (batch == 0) return;
    PubArgs a{};
    a.nlimbs = nlimbs; a.npoly = npoly; a.mul = mul ? 1 : 0; a.n = p->n; a.batch = batch;
    for (int l = 0; l < nlimbs; ++l) { a.limb[l].q = p->limb[l].q; a.limb[l].mu = p->limb[l].mu; }
    k_ct_public<<<symm_grid(batch * (mul ? npoly : 1) * nlimbs * p->n, p->sm_count), 256, 0, (cudaStream_t)stream>>>(a, c, pub);
    ab_count_launch();
    ab_check_launch("ct_public");
}
extern "C" int ab_ct_add_public_dev(const ab_plan *p, int64_t *c, const int64_t *pub, int64_t batch, int32_t npoly, int32_t nlimbs, void *stream) {
    return ab_guard([&] { ct_public(p, c, pub, batch, npoly, nlimbs, false, stream); });
}
extern "C" int ab_ct_mul_public_dev(const ab_plan *p, int64_t *c, const int64_t *pub, int64_t batch, int32_t npoly, int32_t nlimbs, void *stream) {
    return ab_guard([&] { ct_public(p, c, pub, batch, npoly, nlimbs, true, stream); });
}

// ------------------------------------------------------------------------------------ plaintext -> Z_q tuples
struct LpArgs {
    uint32_t q[AB_MAX_LIMBS];
    int32_t nlimbs, pad;
    int64_t n, batch, pmod;
};
__global__ void __launch_bounds__(256) k_lift_plain(const LpArgs a, int64_t *__restrict__ out, const int64_t *__restrict__ in) {
    const int64_t total = a.batch * a.n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / a.n, j = i % a.n;
        int64_t v = in[i] % a.pmod;
        if (v < 0) v += a.pmod;
        if (2 * v >= a.pmod) v -= a.pmod;             // Lol `lift`: representative in [-p/2, p/2)
        for (int l = 0; l < a.nlimbs; ++l) {
            int64_t r = v % (int64_t)a.q[l];
            if (r < 0) r += a.q[l];
            out[(b * a.nlimbs + l) * a.n + j] = r;
        }
    }
}
extern "C" int ab_lift_plain_dev(const ab_plan *p, int64_t *out, const int64_t *in, int64_t batch, int32_t nlimbs, int64_t pmod, void *stream) {
    return ab_guard([&] {
        AB_REQUIRE(p && ((out && in) || batch == 0), AB_ERR_ARG, "null argument");
        AB_REQUIRE(batch >= 0 && nlimbs >= 1 && nlimbs <= p->L && pmod >= 2 && pmod < (1ll << 62), AB_ERR_ARG, "bad batch, limb count or plaintext modulus");
        if (batch == 0) return;
        LpArgs a{};
        a.nlimbs = nlimbs; a.n = p->n; a.batch = batch; a.pmod = pmod;
        for (int l = 0; l < nlimbs; ++l) a.q[l] = p->limb[l].q;
        k_lift_plain<<<symm_grid(batch * p->n, p->sm_count), 256, 0, (cudaStream_t)stream>>>(a, out, in);
        ab_count_launch();
        ab_check_launch("lift_plain");
    });
}

// ------------------------------------------------------------------------------------ centred lift over the RNS product
// Garner: x = v_0 + v_1 q_0 + v_2 q_0 q_1 + ...,  v_i = (r_i - (v_0 + v_1 q_0 + ...)) (q_0 ... q_(i-1))^-1 mod q_i.
// x > (Q-1)/2 is decided digit by digit against the mixed-radix digits of (Q-1)/2 (Q odd), most significant first.
struct RlArgs {
    uint32_t q[AB_MAX_LIMBS];
    uint64_t mu[AB_MAX_LIMBS];
    uint32_t pinv[AB_MAX_LIMBS];                   // (q_0 ... q_(i-1))^-1 mod q_i
    uint32_t pmodq[AB_MAX_LIMBS][AB_MAX_LIMBS];    // [i][k] = q_0 ... q_(k-1) mod q_i   (k < i)
    uint32_t half[AB_MAX_LIMBS];                   // mixed-radix digits of (Q - 1) / 2
    uint64_t pw_p[AB_MAX_LIMBS];                   // q_0 ... q_(i-1) mod pmod
    double pw_d[AB_MAX_LIMBS];                     // q_0 ... q_(i-1) / Q
    uint64_t Q_p;                                  // Q mod pmod
    uint64_t pmod;
    int32_t nlimbs, gx;                            // gx: CTAs per ring element (the batch index lives in blockIdx.x / gx: no 65535 cap)
    int64_t n, batch;
};
__global__ void __launch_bounds__(256) k_rns_lift_reduce(const RlArgs a, int64_t *__restrict__ out, double *__restrict__ rate, const int64_t *__restrict__ in) {
    __shared__ double smax[256];
    const int64_t b = blockIdx.x / a.gx;
    double mx = 0.0;
    for (int64_t j = (int64_t)(blockIdx.x % a.gx) * blockDim.x + threadIdx.x; j < a.n; j += (int64_t)a.gx * blockDim.x) {
        uint32_t v[AB_MAX_LIMBS];
        const int64_t *ip = in + b * a.nlimbs * a.n + j;
#pragma unroll
        for (int i = 0; i < AB_MAX_LIMBS; ++i) {
            if (i >= a.nlimbs) break;
            const uint32_t q = a.q[i];
            uint32_t partial = 0;               // (v_0 + v_1 q_0 + ...) mod q_i
            for (int k = 0; k < i; ++k) partial = ab_addmod(partial, ab_mulmod(v[k] % q, a.pmodq[i][k], q, a.mu[i]), q);
            const uint32_t d = ab_submod((uint32_t)__ldcs(ip + (int64_t)i * a.n), partial, q);
            v[i] = i ? ab_mulmod(d, a.pinv[i], q, a.mu[i]) : d;
        }
        bool neg = false;                       // x > (Q - 1) / 2 ?
        for (int i = a.nlimbs - 1; i >= 0; --i)
            if (v[i] != a.half[i]) { neg = v[i] > a.half[i]; break; }
        // value mod pmod, and |value| / Q as a double (diagnostic: errorRate)
        unsigned __int128 acc = 0;
        for (int i = 0; i < a.nlimbs; ++i) acc += (unsigned __int128)v[i] * a.pw_p[i];
        uint64_t r = (uint64_t)(acc % a.pmod);
        if (neg) {
            r = (r + a.pmod - a.Q_p) % a.pmod;
            // |x| = Q - x = 1 + sum (q_i - 1 - v_i) P_i: take ITS mixed-radix digits (complement plus a carry) so that a small
            // magnitude keeps its leading digits zero -- 1.0 - x/Q would cancel to an absolute error of 1e-16 (error rates are far smaller)
            uint32_t carry = 1;
            for (int i = 0; i < a.nlimbs; ++i) {
                uint32_t w = a.q[i] - 1u - v[i] + carry;
                carry = w == a.q[i];
                v[i] = carry ? 0u : w;
            }
        }
        double xd = 0.0;                        // most significant digit last: the small terms are summed first
        for (int i = 0; i < a.nlimbs; ++i) xd += (double)v[i] * a.pw_d[i];
        out[b * a.n + j] = (int64_t)r;
        mx = fmax(mx, xd);
    }
    if (rate) {
        smax[threadIdx.x] = mx;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if ((int)threadIdx.x < s) smax[threadIdx.x] = fmax(smax[threadIdx.x], smax[threadIdx.x + s]);
            __syncthreads();
        }
        // non-negative doubles order like their bit patterns: an integer atomicMax keeps the maximum
        if (threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned long long *>(rate + b), (unsigned long long)__double_as_longlong(smax[0]));
    }
}
extern "C" int ab_rns_lift_reduce_dev(const ab_plan *p, int64_t *out, double *maxrate, const int64_t *in, int64_t batch, int32_t nlimbs, int64_t pmod, void *stream) {
    return ab_guard([&] {
        AB_REQUIRE(p && ((out && in) || batch == 0), AB_ERR_ARG, "null argument");
        AB_REQUIRE(batch >= 0 && nlimbs >= 1 && nlimbs <= p->L && pmod >= 2 && pmod < (1ll << 62), AB_ERR_ARG, "bad batch, limb count or plaintext modulus");
        if (batch == 0) return;
        RlArgs a{};
        a.nlimbs = nlimbs; a.n = p->n; a.batch = batch; a.pmod = (uint64_t)pmod;
        // all tables from exact integer arithmetic (Q < 2^(31 * 8): a small schoolbook big number)
        std::vector<uint32_t> Q{1};                 // little-endian base 2^32
        auto mul_small = [](std::vector<uint32_t> &x, uint64_t f) {
            uint64_t carry = 0;
            for (auto &w : x) { const uint64_t t = (uint64_t)w * f + carry; w = (uint32_t)t; carry = t >> 32; }
            if (carry) x.push_back((uint32_t)carry);
        };
        auto mod_small = [](const std::vector<uint32_t> &x, uint64_t m) {
            unsigned __int128 r = 0;
            for (size_t i = x.size(); i-- > 0;) r = ((r << 32) | x[i]) % m;
            return (uint64_t)r;
        };
        auto div_small = [](std::vector<uint32_t> &x, uint64_t d) {     // x /= d, returns the remainder
            uint64_t r = 0;
            for (size_t i = x.size(); i-- > 0;) { const uint64_t t = (r << 32) | x[i]; x[i] = (uint32_t)(t / d); r = t % d; }
            while (x.size() > 1 && x.back() == 0) x.pop_back();
            return r;
        };
        for (int i = 0; i < nlimbs; ++i) {
            const uint64_t q = (uint64_t)p->q[i];
            a.q[i] = (uint32_t)q; a.mu[i] = p->limb[i].mu;
            a.pw_p[i] = mod_small(Q, (uint64_t)pmod);
            for (int k = 0; k < i; ++k) {
                uint64_t pr = 1;
                for (int t = 0; t < k; ++t) pr = abh::mulmod(pr, (uint64_t)p->q[t] % q, q);
                a.pmodq[i][k] = (uint32_t)pr;
            }
            const uint64_t Pi = mod_small(Q, q);
            a.pinv[i] = i ? (uint32_t)abh::invmod(Pi, q) : 1u;
            mul_small(Q, q);
        }
        a.Q_p = mod_small(Q, (uint64_t)pmod);
        std::vector<uint32_t> H = Q;               // (Q - 1) / 2: Q is odd (odd prime moduli)
        H[0] -= 1;
        div_small(H, 2);
        for (int i = 0; i < nlimbs; ++i) a.half[i] = (uint32_t)div_small(H, (uint64_t)p->q[i]);
        double w = 1.0;
        for (int i = nlimbs - 1; i >= 0; --i) { w /= (double)p->q[i]; a.pw_d[i] = w; }    // q_0...q_(i-1) / Q = 1 / (q_i ... q_(L-1))
        cudaStream_t st = (cudaStream_t)stream;
        if (maxrate) AB_CHECK_CUDA(cudaMemsetAsync(maxrate, 0, (size_t)batch * sizeof(double), st));
        const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>((p->n + 255) / 256, 64));
        AB_REQUIRE(batch * (int64_t)gx < (1ll << 31), AB_ERR_ARG, "batch too large for one launch");
        a.gx = (int32_t)gx;
        k_rns_lift_reduce<<<(unsigned)(batch * gx), 256, 0, st>>>(a, out, maxrate, in);
        ab_count_launch();
        ab_check_launch("rns_lift_reduce");
    });
}
