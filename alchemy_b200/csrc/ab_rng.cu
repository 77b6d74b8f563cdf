// ab_rng.cu -- counter-based randomness on the device for key / hint generation:
// ALCHEMY draws a uniform c1 per encryption and per hint row and a Gaussian error per hint row
// (Crypto/Alchemy/Interpreter/KeysHints.hs:96,113,115-129 via Lol's getRandom / tGaussianDec [RECALL]).
// Philox4x32-10 (Salmon et al., SC'11) keyed by the caller's 64-bit seed; the counter is the element's position
// (word 0-1), a caller-chosen stream id (word 2) and a draw index (word 3), so every output is a pure function of
// (seed, stream, position): reproducible on any grid, and restated bit for bit on the CPU by the test suite.
//   * uniform residues:  v = floor(r * q / 2^64) for a 64-bit r  (bias < 2^-33 for q < 2^31)
//   * discrete Gaussian: inversion against a cumulative table T[i] = floor(2^64 * P(|X| <= i)) built by the
//     caller in double precision; sign from a separate random bit.  Integer comparisons only: no device
//     floating point enters the sample, so parity with the CPU restatement is exact.
// Deviation from upstream, stated: Lol's tGaussianDec is spherical in the canonical embedding (correlated
// decoding-basis coefficients); here the decoding-basis coefficients are i.i.d.
#include "ab_internal.h"

struct Philox { uint32_t c[4]; };
__device__ __forceinline__ Philox philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return Philox{{c0, c1, c2, c3}};
}

struct UnArgs {
    uint32_t q[AB_MAX_LIMBS];
    int32_t nlimbs, pad;
    int64_t n, batch;
    uint64_t seed;
    uint32_t stream, pad2;
};
__global__ void __launch_bounds__(256) k_uniform(const UnArgs a, int64_t *__restrict__ out) {
    const int64_t total = a.batch * a.nlimbs * a.n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int l = (int)((i / a.n) % a.nlimbs);
        const Philox r = philox4x32_10((uint32_t)i, (uint32_t)((uint64_t)i >> 32), a.stream, 0u, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
        const uint64_t w = ((uint64_t)r.c[0] << 32) | r.c[1];
        out[i] = (int64_t)__umul64hi(w, (uint64_t)a.q[l]);
    }
}
extern "C" int ab_uniform_dev(const ab_plan *p, int64_t *out, int64_t batch, int32_t nlimbs, uint64_t seed, uint32_t stream_id, void *stream) {
    return ab_guard([&] {
        AB_REQUIRE(p && (out || batch == 0), AB_ERR_ARG, "null argument");
        AB_REQUIRE(batch >= 0 && nlimbs >= 1 && nlimbs <= p->L, AB_ERR_ARG, "bad batch or limb count");
        if (batch == 0) return;
        UnArgs a{};
        a.nlimbs = nlimbs; a.n = p->n; a.batch = batch; a.seed = seed; a.stream = stream_id;
        for (int l = 0; l < nlimbs; ++l) a.q[l] = p->limb[l].q;
        const int64_t work = batch * nlimbs * p->n;
        const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((work + 255) / 256, (int64_t)p->sm_count * 16));
        k_uniform<<<grid, 256, 0, (cudaStream_t)stream>>>(a, out);
        ab_count_launch();
        ab_check_launch("uniform");
    });
}

#define AB_CDT_MAX 128
struct GaArgsR {
    uint64_t cdt[AB_CDT_MAX];
    int32_t len, pad;
    int64_t n, batch;
    uint64_t seed;
    uint32_t stream, pad2;
};
__global__ void __launch_bounds__(256) k_gaussian(const __grid_constant__ GaArgsR a, int64_t *__restrict__ out) {
    const int64_t total = a.batch * a.n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const Philox r = philox4x32_10((uint32_t)i, (uint32_t)((uint64_t)i >> 32), a.stream, 1u, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
        const uint64_t u = ((uint64_t)r.c[0] << 32) | r.c[1];
        int mag = 0;                              // number of table entries below or equal to u
        for (int k = 0; k < a.len; ++k) mag += (a.cdt[k] <= u) ? 1 : 0;
        out[i] = (r.c[2] & 1u) ? -(int64_t)mag : (int64_t)mag;
    }
}
extern "C" int ab_gaussian_dev(const ab_plan *p, int64_t *out, int64_t batch, const uint64_t *cdt, int32_t cdt_len, uint64_t seed, uint32_t stream_id, void *stream) {
    return ab_guard([&] {
        AB_REQUIRE(p && (out || batch == 0) && cdt, AB_ERR_ARG, "null argument");
        AB_REQUIRE(batch >= 0 && cdt_len >= 1 && cdt_len <= AB_CDT_MAX, AB_ERR_ARG, "bad batch or table length");
        if (batch == 0) return;
        GaArgsR a{};
        a.len = cdt_len; a.n = p->n; a.batch = batch; a.seed = seed; a.stream = stream_id;
        for (int k = 0; k < cdt_len; ++k) {
            AB_REQUIRE(k == 0 || cdt[k] >= cdt[k - 1], AB_ERR_ARG, "cumulative table must be non-decreasing");
            a.cdt[k] = cdt[k];
        }
        const int64_t work = batch * p->n;
        const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((work + 255) / 256, (int64_t)p->sm_count * 16));
        k_gaussian<<<grid, 256, 0, (cudaStream_t)stream>>>(a, out);
        ab_count_launch();
        ab_check_launch("gaussian");
    });
}
