// ab_elementwise.cu -- streaming kernels: RNS pointwise ops (lol-cpp mulRq, zipWithT), the
// per-prime L / G operators (tensorLRq, tensorGPowRq, ... [RECALL]), index gathers for
// twace/embed/coeffs (done with backpermute in Lol's Haskell), gadget decomposition and RNS
// rescaling (Lol Gadget / RescaleCyc, reached from Eval.hs:130,133).
// All are HBM-bound: one coalesced read and one coalesced write of each int64 residue.
#include "ab_internal.h"
#include "ab_modarith.cuh"
#include "ab_kernels.h"

static inline unsigned grid_for(int64_t work, int threads, int sm_count) {
    int64_t g = (work + threads - 1) / threads;
    int64_t cap = (int64_t)sm_count * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

// ------------------------------------------------------------------------------------ pointwise
struct EwArgs {
    struct {
        uint32_t q, s, s_shoup, pad;
        uint64_t mu;
        const uint2 *vec;    // per-coefficient constants (g or g^-1 at every CRT row)
    } limb[AB_MAX_LIMBS];
    int32_t nlimbs, op, layout;
    int64_t n, total, xmod;  // xmod = elements of x before it repeats (broadcast)
};

__global__ void __launch_bounds__(256) k_elementwise(const EwArgs a, int64_t *__restrict__ y, const int64_t *__restrict__ x) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.total; i += (int64_t)gridDim.x * blockDim.x) {
        int limb;
        int64_t j;
        if (a.layout == AB_LAYOUT_PLANAR) { limb = (int)((i / a.n) % a.nlimbs); j = i % a.n; }
        else { limb = (int)(i % a.nlimbs); j = (i / a.nlimbs) % a.n; }
        const uint32_t q = a.limb[limb].q;
        uint32_t v = (uint32_t)y[i], r;
        switch (a.op) {
            case AB_EW_MUL: r = ab_mulmod(v, (uint32_t)x[i % a.xmod], q, a.limb[limb].mu); break;
            case AB_EW_ADD: r = ab_addmod(v, (uint32_t)x[i % a.xmod], q); break;
            case AB_EW_SUB: r = ab_submod(v, (uint32_t)x[i % a.xmod], q); break;
            case AB_EW_NEG: r = v ? q - v : 0; break;
            case AB_EW_SCALAR: r = ab_shoup(v, a.limb[limb].s, a.limb[limb].s_shoup, q); break;
            default: r = ab_shoup(v, __ldg(a.limb[limb].vec + j), q); break;   // AB_EW_GCRT / AB_EW_GINVCRT
        }
        y[i] = (int64_t)r;
    }
}

void ab_launch_elementwise(const ab_plan *p, int op, const AbView &v, const int64_t *x, int64_t x_batch,
                           const int64_t *scalars_host, cudaStream_t st) {
    if (v.batch == 0) return;
    EwArgs a{};
    a.nlimbs = v.nlimbs; a.op = op; a.layout = v.layout; a.n = p->n;
    a.total = v.batch * v.nlimbs * p->n;
    a.xmod = (x_batch == 1 ? 1 : v.batch) * v.nlimbs * p->n;
    if (op == AB_EW_MUL || op == AB_EW_ADD || op == AB_EW_SUB) {
        AB_REQUIRE(x != nullptr, AB_ERR_ARG, "null operand");
        AB_REQUIRE(x_batch == 1 || x_batch == v.batch, AB_ERR_ARG, "operand batch must be 1 or equal to the batch");
    }
    if (op == AB_EW_GCRT || op == AB_EW_GINVCRT) AB_REQUIRE(p->has_crt, AB_ERR_NO_ROOT, "plan has no CRT tables");
    for (int l = 0; l < v.nlimbs; ++l) {
        const AbLimbDev &ld = p->limb[v.limb0 + l];
        a.limb[l].q = ld.q; a.limb[l].mu = ld.mu;
        a.limb[l].vec = reinterpret_cast<const uint2 *>(op == AB_EW_GCRT ? ld.g_crt : ld.ginv_crt);
        if (op == AB_EW_SCALAR) {
            AB_REQUIRE(scalars_host != nullptr, AB_ERR_ARG, "null scalars");
            int64_t s = scalars_host[l] % (int64_t)ld.q;
            if (s < 0) s += ld.q;
            a.limb[l].s = (uint32_t)s; a.limb[l].s_shoup = abh::shoup((uint64_t)s, ld.q);
        }
    }
    k_elementwise<<<grid_for(a.total, 256, p->sm_count), 256, 0, st>>>(a, v.y, x);
    ab_count_launch();
    ab_check_launch("elementwise");
}

// ------------------------------------------------------------------------------------ L and G along one prime axis
struct PrArgs {
    struct { uint32_t q, pinv, pinv_s, pad; } limb[AB_MAX_LIMBS];
    int32_t nlimbs, op, d;        // d = p - 1
    int64_t n, rts, nvec, total;  // nvec = vectors per ring element = n / d
    int64_t estride, lstride, bstride;
};

__global__ void __launch_bounds__(256) k_prime_op(const PrArgs a, int64_t *__restrict__ y) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t vec = i % a.nvec, row = i / a.nvec;
        const int limb = (int)(row % a.nlimbs);
        const int64_t b = row / a.nlimbs;
        const uint32_t q = a.limb[limb].q;
        const int64_t t = vec / a.rts, r = vec % a.rts;
        int64_t *e = y + b * a.bstride + limb * a.lstride + (t * a.d * a.rts + r) * a.estride;
        const int64_t es = a.rts * a.estride;
        const int d = a.d;
        switch (a.op) {
            case AB_PR_L: {
                uint32_t acc = 0;
                for (int k = 0; k < d; ++k) { acc = ab_addmod(acc, (uint32_t)e[k * es], q); e[k * es] = acc; }
            } break;
            case AB_PR_LINV: {
                uint32_t prev = 0;
                for (int k = 0; k < d; ++k) { uint32_t c = (uint32_t)e[k * es]; e[k * es] = ab_submod(c, prev, q); prev = c; }
            } break;
            case AB_PR_GPOW: {
                const uint32_t last = (uint32_t)e[(d - 1) * es];
                uint32_t prev = 0;
                for (int k = 0; k < d; ++k) {
                    uint32_t c = (uint32_t)e[k * es];
                    e[k * es] = ab_addmod(ab_submod(c, prev, q), last, q);
                    prev = c;
                }
            } break;
            case AB_PR_GINVPOW: {
                uint32_t sum = 0;
                for (int k = 0; k < d; ++k) sum = ab_addmod(sum, (uint32_t)e[k * es], q);
                const uint32_t tt = ab_shoup(sum, a.limb[limb].pinv, a.limb[limb].pinv_s, q);
                uint32_t pre = 0, jt = 0;
                for (int k = 0; k < d; ++k) {
                    pre = ab_addmod(pre, (uint32_t)e[k * es], q);
                    jt = ab_addmod(jt, tt, q);
                    e[k * es] = ab_submod(pre, jt, q);
                }
            } break;
            case AB_PR_GDEC: {
                uint32_t sum = 0;
                for (int k = 0; k < d; ++k) sum = ab_addmod(sum, (uint32_t)e[k * es], q);
                uint32_t prev = 0;
                for (int k = 0; k < d; ++k) {
                    uint32_t c = (uint32_t)e[k * es];
                    e[k * es] = (k == 0) ? ab_addmod(c, sum, q) : ab_submod(c, prev, q);
                    prev = c;
                }
            } break;
            default: {   // AB_PR_GINVDEC
                uint32_t c = 0, sc = 0;
                for (int k = 1; k < d; ++k) { c = ab_addmod(c, (uint32_t)e[k * es], q); sc = ab_addmod(sc, c, q); }
                const uint32_t a0 = ab_shoup(ab_submod((uint32_t)e[0], sc, q), a.limb[limb].pinv, a.limb[limb].pinv_s, q);
                e[0] = a0;
                c = 0;
                for (int k = 1; k < d; ++k) { c = ab_addmod(c, (uint32_t)e[k * es], q); e[k * es] = ab_addmod(a0, c, q); }
            } break;
        }
    }
}

void ab_launch_prime_op(const ab_plan *p, int op, const AbView &v, cudaStream_t st) {
    if (v.batch == 0) return;
    const bool is_g = (op != AB_PR_L && op != AB_PR_LINV);
    for (size_t ai = 0; ai < p->axes.size(); ++ai) {
        const AbAxis &ax = p->axes[ai];
        if (ax.p == 2) continue;                 // L_2 and G_2 are the identity (g runs over odd primes)
        if (op == AB_PR_GINVPOW || op == AB_PR_GINVDEC)
            for (int l = 0; l < v.nlimbs; ++l)
                AB_REQUIRE(p->limb[v.limb0 + l].pinv[ai] != 0, AB_ERR_UNSUPPORTED, "divG: a prime of m is not invertible modulo this modulus");
        (void)is_g;
        PrArgs a{};
        a.nlimbs = v.nlimbs; a.op = op; a.d = ax.p - 1; a.n = p->n; a.rts = ax.rts;
        a.nvec = p->n / a.d;
        a.total = v.batch * v.nlimbs * a.nvec;
        if (v.layout == AB_LAYOUT_PLANAR) { a.estride = 1; a.lstride = p->n; a.bstride = p->n * v.nlimbs; }
        else { a.estride = v.nlimbs; a.lstride = 1; a.bstride = p->n * v.nlimbs; }
        for (int l = 0; l < v.nlimbs; ++l) {
            const AbLimbDev &ld = p->limb[v.limb0 + l];
            a.limb[l].q = ld.q; a.limb[l].pinv = ld.pinv[ai]; a.limb[l].pinv_s = ld.pinv_s[ai];
        }
        k_prime_op<<<grid_for(a.total, 256, p->sm_count), 256, 0, st>>>(a, v.y);
        ab_count_launch();
        ab_check_launch("prime_op");
    }
}

// ------------------------------------------------------------------------------------ index maps
// out[row][j] = sgn * in[in_row(row)][idx[tab_row(row)][j]];  rows are (batch, [k,] limb) tuples.
//   mode 0: plain gather, rows map 1:1            (twacePowDec, embedPow, embedCRT, embedDec)
//   mode 1: coeffs    out row (b,k,l) <- in row (b,l), table row k
//   mode 2: uncoeffs  out row (b,l)[idx[k][j]] <- in row (b,k,l)[j]   (scatter through a permutation)
struct GaArgs {
    uint32_t q[AB_MAX_LIMBS];
    int32_t nlimbs, mode;
    int64_t n_out, n_in, rows, rel;
    const int32_t *idx;
    const int8_t *sgn;
};
__global__ void __launch_bounds__(256) k_gather(const GaArgs a, int64_t *__restrict__ out, const int64_t *__restrict__ in) {
    const int64_t total = a.rows * (a.mode == 2 ? a.n_in : a.n_out);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        if (a.mode == 0) {
            const int64_t row = i / a.n_out, j = i % a.n_out;
            const int32_t s = a.idx[j];
            int64_t v = 0;
            if (s >= 0) {
                v = in[row * a.n_in + s];
                if (a.sgn && a.sgn[j] < 0 && v) v = (int64_t)a.q[row % a.nlimbs] - v;
            }
            out[i] = v;
        } else if (a.mode == 1) {
            const int64_t row = i / a.n_out, j = i % a.n_out;          // row = (b*rel + k)*L + l
            const int64_t l = row % a.nlimbs, k = (row / a.nlimbs) % a.rel, b = row / (a.nlimbs * a.rel);
            out[i] = in[(b * a.nlimbs + l) * a.n_in + a.idx[k * a.n_out + j]];
        } else {
            const int64_t row = i / a.n_in, j = i % a.n_in;            // in row = (b*rel + k)*L + l
            const int64_t l = row % a.nlimbs, k = (row / a.nlimbs) % a.rel, b = row / (a.nlimbs * a.rel);
            out[(b * a.nlimbs + l) * a.n_out + a.idx[k * a.n_in + j]] = in[i];
        }
    }
}
void ab_launch_index_map(int mode, int64_t *out, const int64_t *in, const int32_t *idx, const int8_t *sgn, const ab_plan *p,
                         int64_t rows, int64_t n_out, int64_t n_in, int64_t rel, int32_t nlimbs, cudaStream_t st) {
    if (rows == 0) return;
    GaArgs a{};
    a.nlimbs = nlimbs; a.mode = mode; a.n_out = n_out; a.n_in = n_in; a.rows = rows; a.rel = rel; a.idx = idx; a.sgn = sgn;
    for (int l = 0; l < nlimbs; ++l) a.q[l] = p->limb[l].q;
    const int64_t total = rows * (mode == 2 ? n_in : n_out);
    k_gather<<<grid_for(total, 256, p->sm_count), 256, 0, st>>>(a, out, in);
    ab_count_launch();
    ab_check_launch("index_map");
}

// twaceCRT: out[row][i] = sum_k w[limb][above[i][k]] * in[row][above[i][k]]
struct TwArgs {
    uint32_t q[AB_MAX_LIMBS];
    int32_t nlimbs;
    int64_t n_lo, n_hi, rows, rel;
    const int32_t *above;
    const uint2 *w;     // [L][n_hi]
};
__global__ void __launch_bounds__(256) k_twace_crt(const TwArgs a, int64_t *__restrict__ out, const int64_t *__restrict__ in) {
    const int64_t total = a.rows * a.n_lo;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / a.n_lo, j = i % a.n_lo;
        const int limb = (int)(row % a.nlimbs);
        const uint32_t q = a.q[limb];
        const uint2 *w = a.w + (int64_t)limb * a.n_hi;
        uint32_t acc = 0;
        for (int64_t k = 0; k < a.rel; ++k) {
            const int32_t s = a.above[j * a.rel + k];
            acc = ab_addmod(acc, ab_shoup((uint32_t)in[row * a.n_hi + s], __ldg(w + s), q), q);
        }
        out[i] = (int64_t)acc;
    }
}
void ab_launch_twace_crt(const ab_ext *e, int64_t *out, const int64_t *in, int64_t batch, int32_t nlimbs, cudaStream_t st) {
    if (batch == 0) return;
    TwArgs a{};
    a.nlimbs = nlimbs; a.n_lo = e->n_lo; a.n_hi = e->n_hi; a.rows = batch * nlimbs; a.rel = e->rel;
    a.above = e->d_above_crt; a.w = reinterpret_cast<const uint2 *>(e->d_tw_weight);
    for (int l = 0; l < nlimbs; ++l) a.q[l] = e->lo->limb[l].q;
    k_twace_crt<<<grid_for(a.rows * a.n_lo, 256, e->lo->sm_count), 256, 0, st>>>(a, out, in);
    ab_count_launch();
    ab_check_launch("twace_crt");
}

// ------------------------------------------------------------------------------------ layout
__global__ void __launch_bounds__(256) k_layout(int64_t *__restrict__ out, const int64_t *__restrict__ in, int64_t batch, int L, int64_t n,
                                                int to_planar) {
    const int64_t total = batch * L * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        // i indexes the planar side: (b, l, j)
        const int64_t j = i % n, l = (i / n) % L, b = i / (n * L);
        const int64_t k = (b * n + j) * L + l;
        if (to_planar) out[i] = in[k]; else out[k] = in[i];
    }
}
void ab_launch_layout(int64_t *out, const int64_t *in, int64_t batch, int32_t L, int64_t n, bool to_planar, cudaStream_t st) {
    if (batch == 0) return;
    k_layout<<<grid_for(batch * L * n, 256, 148), 256, 0, st>>>(out, in, batch, L, n, to_planar ? 1 : 0);
    ab_count_launch();
    ab_check_launch("layout");
}

// ------------------------------------------------------------------------------------ gadget decomposition
struct DcArgs {
    uint32_t q[AB_MAX_LIMBS];
    int32_t digits[AB_MAX_LIMBS];     // digits contributed by each input limb
    int32_t first[AB_MAX_LIMBS];      // index of the first of them
    int32_t nlimbs, out_limbs, base, ndigits;
    int64_t n, batch;
};
// in[b][l][j] -> out[b][d][l'][j]
__global__ void __launch_bounds__(256) k_decompose(const DcArgs a, int64_t *__restrict__ out, const int64_t *__restrict__ in) {
    const int64_t total = a.batch * a.nlimbs * a.n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = i % a.n, b = i / (a.n * a.nlimbs);
        const int l = (int)((i / a.n) % a.nlimbs);
        int32_t v = ab_center((uint32_t)in[i], a.q[l]);
        int64_t *o = out + ((b * a.ndigits + a.first[l]) * a.out_limbs) * a.n + j;
        const int nd = a.digits[l];
        for (int d = 0; d < nd; ++d) {
            int32_t r;
            if (a.base == 0 || d == nd - 1) { r = v; }
            else {   // Lol divModCent: floor division, remainder moved into [-b/2, b/2)
                int32_t qq = v / a.base, rr = v - qq * a.base;
                if (rr < 0) { rr += a.base; qq -= 1; }
                if (2 * rr >= a.base) { rr -= a.base; qq += 1; }
                r = rr; v = qq;
            }
            for (int lp = 0; lp < a.out_limbs; ++lp) o[(int64_t)(d * a.out_limbs + lp) * a.n] = (int64_t)ab_reduce_signed(r, a.q[lp]);
        }
    }
}
int32_t ab_gadget_digits(int64_t q, int32_t base) {
    if (base == 0) return 1;
    int32_t k = 0;
    int64_t v = 1;
    while (v < q) { v *= base; ++k; }      // Lol logCeil: smallest k with base^k >= q
    return k;
}
void ab_launch_decompose(const ab_plan *p, int64_t *out, const int64_t *in, int64_t batch, int32_t nlimbs, int32_t out_limbs,
                         int32_t base, cudaStream_t st) {
    if (batch == 0) return;
    DcArgs a{};
    a.nlimbs = nlimbs; a.out_limbs = out_limbs; a.base = base; a.n = p->n; a.batch = batch;
    int32_t nd = 0;
    for (int l = 0; l < nlimbs; ++l) {
        a.digits[l] = ab_gadget_digits(p->q[l], base);
        a.first[l] = nd;
        nd += a.digits[l];
    }
    for (int l = 0; l < std::max(nlimbs, out_limbs); ++l) a.q[l] = p->limb[l].q;
    a.ndigits = nd;
    k_decompose<<<grid_for(batch * nlimbs * p->n, 256, p->sm_count), 256, 0, st>>>(a, out, in);
    ab_count_launch();
    ab_check_launch("decompose");
}

// ------------------------------------------------------------------------------------ RNS rescale
struct RsArgs {
    uint32_t q[AB_MAX_LIMBS];          // kept limbs, in output order
    uint32_t qdinv[AB_MAX_LIMBS], qdinv_s[AB_MAX_LIMBS];
    int32_t src[AB_MAX_LIMBS];         // input limb index of each kept limb
    uint32_t qd;
    int32_t nin, nout, drop, mode;     // mode 0: in holds residues; 1: `lifted` holds centred-lift CRTs (fused modSwitch)
    int64_t n, batch;
};
// mode 0 (Pow/Dec basis): out[b][k][j] = (in[b][src k][j] - center(in[b][drop][j])) * qd^-1 mod q_k
__global__ void __launch_bounds__(256) k_rescale_drop(const RsArgs a, int64_t *__restrict__ out, const int64_t *__restrict__ in) {
    const int64_t total = a.batch * a.n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = i % a.n, b = i / a.n;
        const int32_t c = ab_center((uint32_t)in[(b * a.nin + a.drop) * a.n + j], a.qd);
        uint32_t vals[AB_MAX_LIMBS];
#pragma unroll
        for (int k = 0; k < AB_MAX_LIMBS; ++k)
            if (k < a.nout) vals[k] = (uint32_t)in[(b * a.nin + a.src[k]) * a.n + j];
#pragma unroll
        for (int k = 0; k < AB_MAX_LIMBS; ++k)
            if (k < a.nout) {
                const uint32_t q = a.q[k];
                const uint32_t d = ab_submod(vals[k], ab_reduce_signed(c, q), q);
                out[(b * a.nout + k) * a.n + j] = (int64_t)ab_shoup(d, a.qdinv[k], a.qdinv_s[k], q);
            }
    }
}
void ab_launch_rescale_drop(const ab_plan *p, int64_t *out, const int64_t *in, int64_t batch, int32_t nlimbs, int32_t drop, cudaStream_t st) {
    if (batch == 0) return;
    RsArgs a{};
    a.nin = nlimbs; a.nout = nlimbs - 1; a.drop = drop; a.n = p->n; a.batch = batch; a.qd = p->limb[drop].q;
    int k = 0;
    for (int l = 0; l < nlimbs; ++l) {
        if (l == drop) continue;
        const uint64_t q = p->q[l];
        a.q[k] = (uint32_t)q; a.src[k] = l;
        const uint64_t inv = abh::invmod((uint64_t)p->q[drop] % q, q);
        a.qdinv[k] = (uint32_t)inv; a.qdinv_s[k] = abh::shoup(inv, q);
        ++k;
    }
    k_rescale_drop<<<grid_for(batch * p->n, 256, p->sm_count), 256, 0, st>>>(a, out, in);
    ab_count_launch();
    ab_check_launch("rescale_drop");
}

// ------------------------------------------------------------------------------------ pieces of the fused SHE ops
// centred lift of one limb reduced into `nout` other moduli: out[b][k][j] = center(in[b][j]) mod q_k
struct LrArgs {
    uint32_t q[AB_MAX_LIMBS];
    uint32_t qsrc;
    int32_t nout;
    int64_t n, batch, in_bstride, out_bstride;
};
__global__ void __launch_bounds__(256) k_lift_reduce(const LrArgs a, int64_t *__restrict__ out, const int64_t *__restrict__ in) {
    const int64_t total = a.batch * a.n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = i % a.n, b = i / a.n;
        const int32_t c = ab_center((uint32_t)in[b * a.in_bstride + j], a.qsrc);
        for (int k = 0; k < a.nout; ++k) out[b * a.out_bstride + k * a.n + j] = (int64_t)ab_reduce_signed(c, a.q[k]);
    }
}
void ab_launch_lift_reduce(const ab_plan *p, int64_t *out, int64_t out_bstride, const int64_t *in, int64_t in_bstride, int64_t batch,
                           int32_t src_limb, const int32_t *dst_limbs, int32_t nout, cudaStream_t st) {
    if (batch == 0 || nout == 0) return;
    LrArgs a{};
    a.qsrc = p->limb[src_limb].q; a.nout = nout; a.n = p->n; a.batch = batch; a.in_bstride = in_bstride; a.out_bstride = out_bstride;
    for (int k = 0; k < nout; ++k) a.q[k] = p->limb[dst_limbs[k]].q;
    k_lift_reduce<<<grid_for(batch * p->n, 256, p->sm_count), 256, 0, st>>>(a, out, in);
    ab_count_launch();
    ab_check_launch("lift_reduce");
}

// out[b][k][j] = (in[b][k][j] - corr[b][k][j]) * s_k mod q_k     (k over `nl` limbs; strided batches)
struct SsArgs {
    uint32_t q[AB_MAX_LIMBS], s[AB_MAX_LIMBS], s_s[AB_MAX_LIMBS];
    int32_t nl;
    int64_t n, batch, in_bstride, corr_bstride, out_bstride;
};
__global__ void __launch_bounds__(256) k_sub_scale(const SsArgs a, int64_t *__restrict__ out, const int64_t *__restrict__ in,
                                                   const int64_t *__restrict__ corr) {
    const int64_t per = (int64_t)a.nl * a.n, total = a.batch * per;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / per, r = i % per;
        const int k = (int)(r / a.n);
        const uint32_t q = a.q[k];
        const uint32_t d = ab_submod((uint32_t)in[b * a.in_bstride + r], (uint32_t)corr[b * a.corr_bstride + r], q);
        out[b * a.out_bstride + r] = (int64_t)ab_shoup(d, a.s[k], a.s_s[k], q);
    }
}
void ab_launch_sub_scale(const ab_plan *p, int64_t *out, int64_t out_bstride, const int64_t *in, int64_t in_bstride,
                         const int64_t *corr, int64_t corr_bstride, int64_t batch, int32_t nl, const int64_t *scal, cudaStream_t st) {
    if (batch == 0 || nl == 0) return;
    SsArgs a{};
    a.nl = nl; a.n = p->n; a.batch = batch; a.in_bstride = in_bstride; a.corr_bstride = corr_bstride; a.out_bstride = out_bstride;
    for (int k = 0; k < nl; ++k) { a.q[k] = p->limb[k].q; a.s[k] = (uint32_t)scal[k]; a.s_s[k] = abh::shoup((uint64_t)scal[k], p->q[k]); }
    k_sub_scale<<<grid_for(batch * nl * p->n, 256, p->sm_count), 256, 0, st>>>(a, out, in, corr);
    ab_count_launch();
    ab_check_launch("sub_scale");
}

// SymmSHE (*): out3[b] = g * (a0 b0, a0 b1 + a1 b0, a1 b1), CRT basis
struct CmArgs {
    struct { uint32_t q, up, up_s, pad; uint64_t mu; const uint2 *g; } limb[AB_MAX_LIMBS];
    int32_t nlimbs, nout;      // nout > nlimbs: the product is written at nout limbs, scaled by the appended moduli (rescale up fused)
    int32_t scaled, pad2;      // every output is multiplied by limb[l].up (appended moduli and / or the caller's per-limb scalar)
    int64_t n, batch;
};
__global__ void __launch_bounds__(256) k_ct_mul(const CmArgs a, int64_t *__restrict__ out, const int64_t *__restrict__ x, const int64_t *__restrict__ y) {
    const int64_t per = (int64_t)a.nlimbs * a.n, total = a.batch * per;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / per, r = i % per;
        const int l = (int)(r / a.n);
        const int64_t j = r % a.n;
        const uint32_t q = a.limb[l].q;
        const uint64_t mu = a.limb[l].mu;
        const uint32_t a0 = (uint32_t)x[(b * 2 + 0) * per + r], a1 = (uint32_t)x[(b * 2 + 1) * per + r];
        const uint32_t b0 = (uint32_t)y[(b * 2 + 0) * per + r], b1 = (uint32_t)y[(b * 2 + 1) * per + r];
        const uint2 g = __ldg(a.limb[l].g + j);
        const uint32_t c0 = ab_mulmod(a0, b0, q, mu), c2 = ab_mulmod(a1, b1, q, mu);
        const uint32_t c1 = ab_addmod(ab_mulmod(a0, b1, q, mu), ab_mulmod(a1, b0, q, mu), q);
        uint32_t o0 = ab_shoup(c0, g, q), o1 = ab_shoup(c1, g, q), o2 = ab_shoup(c2, g, q);
        if (a.scaled) {
            o0 = ab_shoup(o0, a.limb[l].up, a.limb[l].up_s, q);
            o1 = ab_shoup(o1, a.limb[l].up, a.limb[l].up_s, q);
            o2 = ab_shoup(o2, a.limb[l].up, a.limb[l].up_s, q);
        }
        const int64_t pero = (int64_t)a.nout * a.n;
        out[(b * 3 + 0) * pero + r] = (int64_t)o0;
        out[(b * 3 + 1) * pero + r] = (int64_t)o1;
        out[(b * 3 + 2) * pero + r] = (int64_t)o2;
    }
    if (a.nout > a.nlimbs) {       // appended limbs are zero
        const int64_t pero = (int64_t)a.nout * a.n, tail = pero - per, ztotal = a.batch * 3 * tail;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ztotal; i += (int64_t)gridDim.x * blockDim.x)
            out[(i / tail) * pero + per + i % tail] = 0;
    }
}
void ab_launch_ct_mul(const ab_plan *p, int64_t *out3, const int64_t *a2, const int64_t *b2, int64_t batch, int32_t nlimbs, cudaStream_t st) {
    ab_launch_ct_mul_up(p, out3, a2, b2, batch, nlimbs, nlimbs, st);
}
// (*) written directly at nout >= nlimbs limbs: modSwitch up (multiply by the appended moduli, zero limbs) in the same pass
void ab_launch_ct_mul_up(const ab_plan *p, int64_t *out3, const int64_t *a2, const int64_t *b2, int64_t batch, int32_t nlimbs, int32_t nout, cudaStream_t st,
                         const int64_t *prescale) {
    if (batch == 0) return;
    CmArgs a{};
    a.nlimbs = nlimbs; a.nout = nout; a.n = p->n; a.batch = batch; a.scaled = (nout > nlimbs || prescale) ? 1 : 0;
    for (int l = 0; l < nlimbs; ++l) {
        a.limb[l].q = p->limb[l].q; a.limb[l].mu = p->limb[l].mu; a.limb[l].g = reinterpret_cast<const uint2 *>(p->limb[l].g_crt);
        const uint64_t q = (uint64_t)p->q[l];
        uint64_t up = 1;
        for (int k = nlimbs; k < nout; ++k) up = abh::mulmod(up, (uint64_t)p->q[k] % q, q);
        if (prescale) up = abh::mulmod(up, (uint64_t)(((prescale[l] % (int64_t)q) + (int64_t)q) % (int64_t)q), q);
        a.limb[l].up = (uint32_t)up; a.limb[l].up_s = abh::shoup(up, q);
    }
    k_ct_mul<<<grid_for(batch * nlimbs * p->n, 256, p->sm_count), 256, 0, st>>>(a, out3, a2, b2);
    ab_count_launch();
    ab_check_launch("ct_mul");
}

// key-switch inner product: acc[b][i][l][j] (+)= sum_d dig[b][d][l][j] * hint[d][i][l][j]   (i in {0,1})
// `add` (may be null) is the (c0,c1) part the knapsack is added to; dig may be strided per batch element.
struct IpArgs {
    struct { uint32_t q, pad; uint64_t mu; } limb[AB_MAX_LIMBS];
    int32_t nlimbs, ndig;
    int64_t n, batch, dig_bstride, add_bstride, out_bstride;
};
__global__ void __launch_bounds__(256) k_hint_ip(const IpArgs a, int64_t *__restrict__ out, const int64_t *__restrict__ dig,
                                                 const int64_t *__restrict__ hint, const int64_t *__restrict__ add, int accumulate) {
    const int64_t per = (int64_t)a.nlimbs * a.n, total = a.batch * per;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / per, r = i % per;
        const int l = (int)(r / a.n);
        const uint32_t q = a.limb[l].q;
        const uint64_t mu = a.limb[l].mu;
        uint32_t s0 = 0, s1 = 0;
        if (add) { s0 = (uint32_t)add[b * a.add_bstride + r]; s1 = (uint32_t)add[b * a.add_bstride + per + r]; }
        if (accumulate) { s0 = ab_addmod(s0, (uint32_t)out[b * a.out_bstride + r], q); s1 = ab_addmod(s1, (uint32_t)out[b * a.out_bstride + per + r], q); }
        for (int d = 0; d < a.ndig; ++d) {
            const uint32_t v = (uint32_t)dig[b * a.dig_bstride + d * per + r];
            s0 = ab_addmod(s0, ab_mulmod(v, (uint32_t)__ldg(hint + (d * 2 + 0) * per + r), q, mu), q);
            s1 = ab_addmod(s1, ab_mulmod(v, (uint32_t)__ldg(hint + (d * 2 + 1) * per + r), q, mu), q);
        }
        out[b * a.out_bstride + r] = (int64_t)s0;
        out[b * a.out_bstride + per + r] = (int64_t)s1;
    }
}
void ab_launch_hint_ip(const ab_plan *p, int64_t *out, int64_t out_bstride, const int64_t *dig, int64_t dig_bstride, const int64_t *hint,
                       const int64_t *add, int64_t add_bstride, bool accumulate, int64_t batch, int32_t nlimbs, int32_t ndig, cudaStream_t st) {
    if (batch == 0) return;
    IpArgs a{};
    a.nlimbs = nlimbs; a.ndig = ndig; a.n = p->n; a.batch = batch; a.dig_bstride = dig_bstride; a.add_bstride = add_bstride; a.out_bstride = out_bstride;
    for (int l = 0; l < nlimbs; ++l) { a.limb[l].q = p->limb[l].q; a.limb[l].mu = p->limb[l].mu; }
    k_hint_ip<<<grid_for(batch * nlimbs * p->n, 256, p->sm_count), 256, 0, st>>>(a, out, dig, hint, add, accumulate ? 1 : 0);
    ab_count_launch();
    ab_check_launch("hint_ip");
}

// Prepared hints: every hint residue h becomes the pair (h 2^32 mod q, h 2^48 mod q), so that the inner product
// over the gadget digits can be accumulated exactly in 64 bits and reduced once (Montgomery) -- the same
// arithmetic as the dense steps of ab_crt_mixed.cu.  rows = digits * 2; out[row][l][j] as uint2.
struct HpArgs {
    struct { uint32_t q, r32, r32_s, r16, r16_s, pad; } limb[AB_MAX_LIMBS];
    int32_t nlimbs;
    int64_t n, rows;
};
__global__ void __launch_bounds__(256) k_hint_prep(const HpArgs a, uint2 *__restrict__ out, const int64_t *__restrict__ hint) {
    const int64_t per = (int64_t)a.nlimbs * a.n, total = a.rows * per;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int l = (int)((i % per) / a.n);
        const uint32_t q = a.limb[l].q;
        const uint32_t h = (uint32_t)hint[i];
        const uint32_t h32 = ab_shoup(h, a.limb[l].r32, a.limb[l].r32_s, q);
        out[i] = make_uint2(h32, ab_shoup(h32, a.limb[l].r16, a.limb[l].r16_s, q));
    }
}
void ab_launch_hint_prep(const ab_plan *p, void *out, const int64_t *hint, int64_t rows, int32_t nlimbs, cudaStream_t st) {
    if (rows == 0) return;
    HpArgs a{};
    a.nlimbs = nlimbs; a.n = p->n; a.rows = rows;
    for (int l = 0; l < nlimbs; ++l) {
        const uint64_t q = p->q[l], r32 = (1ull << 32) % q, r16 = (1ull << 16) % q;
        a.limb[l].q = (uint32_t)q; a.limb[l].r32 = (uint32_t)r32; a.limb[l].r32_s = abh::shoup(r32, q);
        a.limb[l].r16 = (uint32_t)r16; a.limb[l].r16_s = abh::shoup(r16, q);
    }
    k_hint_prep<<<grid_for(rows * nlimbs * p->n, 256, p->sm_count), 256, 0, st>>>(a, reinterpret_cast<uint2 *>(out), hint);
    ab_count_launch();
    ab_check_launch("hint_prep");
}

// Prepared hints of the fused gadget inner product (k_crt_mixed_mac): both components of a digit's hint side by
// side, each pre-multiplied by 2^32:  out[kd][l][j] = (hint[kd][0][l][j] 2^32 mod q_l, hint[kd][1][l][j] 2^32 mod q_l)
__global__ void __launch_bounds__(256) k_hint_prep2(const HpArgs a, uint2 *__restrict__ out, const int64_t *__restrict__ hint) {
    const int64_t per = (int64_t)a.nlimbs * a.n, total = a.rows * per;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t kd = i / per, r = i % per;
        const int l = (int)(r / a.n);
        const uint32_t q = a.limb[l].q;
        const uint32_t h0 = (uint32_t)hint[(2 * kd) * per + r], h1 = (uint32_t)hint[(2 * kd + 1) * per + r];
        out[i] = make_uint2(ab_shoup(h0, a.limb[l].r32, a.limb[l].r32_s, q), ab_shoup(h1, a.limb[l].r32, a.limb[l].r32_s, q));
    }
}
void ab_launch_hint_prep2(const ab_plan *p, void *out, const int64_t *hint, int64_t ndig, int32_t nlimbs, cudaStream_t st) {
    if (ndig == 0) return;
    HpArgs a{};
    a.nlimbs = nlimbs; a.n = p->n; a.rows = ndig;
    for (int l = 0; l < nlimbs; ++l) {
        const uint64_t q = p->q[l], r32 = (1ull << 32) % q;
        a.limb[l].q = (uint32_t)q; a.limb[l].r32 = (uint32_t)r32; a.limb[l].r32_s = abh::shoup(r32, q);
    }
    k_hint_prep2<<<grid_for(ndig * nlimbs * p->n, 256, p->sm_count), 256, 0, st>>>(a, reinterpret_cast<uint2 *>(out), hint);
    ab_count_launch();
    ab_check_launch("hint_prep2");
}

// out[b][i][l][j] = add[b][i][l][j] + sum_d dig[b][d][l][j] * hint[d][i][l][j]   with prepared hints
struct IpmArgs {
    struct { uint32_t q, qinv_neg, wide, pad; uint64_t mu; } limb[AB_MAX_LIMBS];
    int32_t nlimbs, ndig;
    int64_t n, batch, dig_bstride, add_bstride, out_bstride;
};
template <bool WIDE>
__device__ __forceinline__ void hint_ip_elem(const IpmArgs &a, int l, int64_t b, int64_t r, int64_t per, int64_t *__restrict__ out,
                                             const int64_t *__restrict__ dig, const uint2 *__restrict__ hm, const int64_t *__restrict__ add) {
    const uint32_t q = a.limb[l].q, qinv_neg = a.limb[l].qinv_neg;
    uint64_t t0 = 0, t1 = 0;      // running 64-bit sums of full products
    uint64_t s0 = 0, s1 = 0;      // sums of reduced chunks (narrow path)
    const int64_t *dp = dig + b * a.dig_bstride + r;
    const uint2 *hp = hm + r;
    for (int d = 0; d < a.ndig; ++d) {
        const uint32_t x = (uint32_t)__ldcs(dp + (int64_t)d * per);
        const uint2 h0 = __ldg(hp + (int64_t)(2 * d) * per), h1 = __ldg(hp + (int64_t)(2 * d + 1) * per);
        if (WIDE) {     // terms below 2^47: no intermediate reduction for any realistic digit count
            const uint32_t xl = x & 0xffffu, xh = x >> 16;
            t0 += (uint64_t)xl * h0.x + (uint64_t)xh * h0.y;
            t1 += (uint64_t)xl * h1.x + (uint64_t)xh * h1.y;
        } else {        // q < 2^30: products below 2^60, reduce every 8 terms
            t0 += (uint64_t)x * h0.x;
            t1 += (uint64_t)x * h1.x;
            if ((d & 7) == 7) {
                s0 += (t0 + (uint64_t)((uint32_t)t0 * qinv_neg) * q) >> 32; t0 = 0;
                s1 += (t1 + (uint64_t)((uint32_t)t1 * qinv_neg) * q) >> 32; t1 = 0;
            }
        }
    }
    s0 += (t0 + (uint64_t)((uint32_t)t0 * qinv_neg) * q) >> 32;
    s1 += (t1 + (uint64_t)((uint32_t)t1 * qinv_neg) * q) >> 32;
    if (add) { s0 += (uint32_t)add[b * a.add_bstride + r]; s1 += (uint32_t)add[b * a.add_bstride + per + r]; }
    // s < 2^32 * (ndig / 8 + 2): one Barrett reduction
    const uint64_t mu = a.limb[l].mu;
    uint64_t r0 = s0 - __umul64hi(s0, mu) * q, r1 = s1 - __umul64hi(s1, mu) * q;
    if (r0 >= q) r0 -= q;
    if (r1 >= q) r1 -= q;
    out[b * a.out_bstride + r] = (int64_t)r0;
    out[b * a.out_bstride + per + r] = (int64_t)r1;
}
__global__ void __launch_bounds__(256) k_hint_ip_m(const IpmArgs a, int64_t *__restrict__ out, const int64_t *__restrict__ dig,
                                                   const uint2 *__restrict__ hm, const int64_t *__restrict__ add) {
    const int64_t per = (int64_t)a.nlimbs * a.n, total = a.batch * per;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / per, r = i % per;
        const int l = (int)(r / a.n);
        if (a.limb[l].wide) hint_ip_elem<true>(a, l, b, r, per, out, dig, hm, add);
        else hint_ip_elem<false>(a, l, b, r, per, out, dig, hm, add);
    }
}
static uint32_t neg_inv32_ew(uint32_t q) {
    uint32_t x = q;
    for (int i = 0; i < 5; ++i) x *= 2u - q * x;
    return 0u - x;
}
void ab_launch_hint_ip_m(const ab_plan *p, int64_t *out, int64_t out_bstride, const int64_t *dig, int64_t dig_bstride, const void *hm,
                         const int64_t *add, int64_t add_bstride, int64_t batch, int32_t nlimbs, int32_t ndig, cudaStream_t st) {
    if (batch == 0) return;
    IpmArgs a{};
    a.nlimbs = nlimbs; a.ndig = ndig; a.n = p->n; a.batch = batch; a.dig_bstride = dig_bstride; a.add_bstride = add_bstride; a.out_bstride = out_bstride;
    for (int l = 0; l < nlimbs; ++l) {
        a.limb[l].q = p->limb[l].q; a.limb[l].qinv_neg = neg_inv32_ew(p->limb[l].q); a.limb[l].mu = p->limb[l].mu;
        a.limb[l].wide = p->limb[l].q >= (1u << 30);
    }
    k_hint_ip_m<<<grid_for(batch * nlimbs * p->n, 256, p->sm_count), 256, 0, st>>>(a, out, dig, reinterpret_cast<const uint2 *>(hm), add);
    ab_count_launch();
    ab_check_launch("hint_ip_m");
}

// acc[b][l][j] (+)= sum_k x[b][k][l][j] * f[k][l][j]      (evalLin of the tunnel's linear function)
// idx != nullptr: x lives in a sub-ring with n_in rows per limb and is replicated through idx (embedCRT): x[b][k][l][idx[j]]
__global__ void __launch_bounds__(256) k_lin_ip(const IpArgs a, int64_t *__restrict__ out, const int64_t *__restrict__ x, const int64_t *__restrict__ f,
                                                const int32_t *__restrict__ idx, int64_t n_in) {
    const int64_t per = (int64_t)a.nlimbs * a.n, total = a.batch * per, per_in = idx ? (int64_t)a.nlimbs * n_in : per;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / per, r = i % per;
        const int l = (int)(r / a.n);
        const uint32_t q = a.limb[l].q;
        const int64_t rin = idx ? (int64_t)l * n_in + __ldg(idx + (r - (int64_t)l * a.n)) : r;
        uint32_t s = 0;
        for (int k = 0; k < a.ndig; ++k)
            s = ab_addmod(s, ab_mulmod((uint32_t)x[b * a.dig_bstride + k * per_in + rin], (uint32_t)__ldg(f + k * per + r), q, a.limb[l].mu), q);
        out[b * a.out_bstride + r] = (int64_t)s;
    }
}
void ab_launch_lin_ip(const ab_plan *p, int64_t *out, int64_t out_bstride, const int64_t *x, int64_t x_bstride, const int64_t *f,
                      int64_t batch, int32_t nlimbs, int32_t nk, cudaStream_t st) {
    ab_launch_lin_ip_rep(p, out, out_bstride, x, x_bstride, f, nullptr, 0, batch, nlimbs, nk, st);
}
void ab_launch_lin_ip_rep(const ab_plan *p, int64_t *out, int64_t out_bstride, const int64_t *x, int64_t x_bstride, const int64_t *f,
                          const int32_t *idx, int64_t n_in, int64_t batch, int32_t nlimbs, int32_t nk, cudaStream_t st) {
    if (batch == 0) return;
    IpArgs a{};
    a.nlimbs = nlimbs; a.ndig = nk; a.n = p->n; a.batch = batch; a.dig_bstride = x_bstride; a.out_bstride = out_bstride;
    for (int l = 0; l < nlimbs; ++l) { a.limb[l].q = p->limb[l].q; a.limb[l].mu = p->limb[l].mu; }
    k_lin_ip<<<grid_for(batch * nlimbs * p->n, 256, p->sm_count), 256, 0, st>>>(a, out, x, f, idx, n_in);
    ab_count_launch();
    ab_check_launch("lin_ip");
}

// strided row copy: out[b*ob + r] = in[b*ib + r], r < len
__global__ void __launch_bounds__(256) k_copy_rows(int64_t *__restrict__ out, int64_t ob, const int64_t *__restrict__ in, int64_t ib, int64_t len, int64_t batch) {
    const int64_t total = batch * len;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / len, r = i % len;
        out[b * ob + r] = in[b * ib + r];
    }
}
void ab_launch_copy_rows(int64_t *out, int64_t ob, const int64_t *in, int64_t ib, int64_t len, int64_t batch, cudaStream_t st) {
    if (batch == 0 || len == 0) return;
    k_copy_rows<<<grid_for(batch * len, 256, 148), 256, 0, st>>>(out, ob, in, ib, len, batch);
    ab_count_launch();
    ab_check_launch("copy_rows");
}
// out[b*ob + k*n + j] = in[b*ib + k*n + j] * s_k   (scale up by the new moduli), and zero-fill of appended limbs
struct UpArgs {
    uint32_t q[AB_MAX_LIMBS], s[AB_MAX_LIMBS], s_s[AB_MAX_LIMBS];
    int32_t nfrom, nto;
    int64_t n, batch, ib, ob;
};
__global__ void __launch_bounds__(256) k_rescale_up(const UpArgs a, int64_t *__restrict__ out, const int64_t *__restrict__ in) {
    const int64_t per = (int64_t)a.nto * a.n, total = a.batch * per;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / per, r = i % per;
        const int k = (int)(r / a.n);
        out[b * a.ob + r] = (k < a.nfrom) ? (int64_t)ab_shoup((uint32_t)in[b * a.ib + r], a.s[k], a.s_s[k], a.q[k]) : 0;
    }
}
void ab_launch_rescale_up(const ab_plan *p, int64_t *out, int64_t ob, const int64_t *in, int64_t ib, int64_t batch, int32_t nfrom, int32_t nto, cudaStream_t st) {
    if (batch == 0) return;
    UpArgs a{};
    a.nfrom = nfrom; a.nto = nto; a.n = p->n; a.batch = batch; a.ib = ib; a.ob = ob;
    for (int k = 0; k < nfrom; ++k) {
        const uint64_t q = p->q[k];
        uint64_t s = 1;
        for (int l = nfrom; l < nto; ++l) s = abh::mulmod(s, (uint64_t)p->q[l] % q, q);
        a.q[k] = (uint32_t)q; a.s[k] = (uint32_t)s; a.s_s[k] = abh::shoup(s, q);
    }
    k_rescale_up<<<grid_for(batch * nto * p->n, 256, p->sm_count), 256, 0, st>>>(a, out, in);
    ab_count_launch();
    ab_check_launch("rescale_up");
}

// ------------------------------------------------------------------------------------ tunnel helpers
struct TgArgs {
    uint32_t q[AB_MAX_LIMBS];
    int32_t nlimbs;
    int64_t n_out, n_in, K, batch, in_bstride;
    const int32_t *tab;
    const int8_t *sgn;
};
__global__ void __launch_bounds__(256) k_tunnel_gather(const TgArgs a, int64_t *__restrict__ out, const int64_t *__restrict__ in) {
    const int64_t total = a.batch * a.K * a.nlimbs * a.n_out;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = i % a.n_out;
        const int l = (int)((i / a.n_out) % a.nlimbs);
        const int64_t k = (i / (a.n_out * a.nlimbs)) % a.K, b = i / (a.n_out * a.nlimbs * a.K);
        const int32_t s = a.tab[k * a.n_out + j];
        int64_t v = 0;
        if (s >= 0) {
            v = in[b * a.in_bstride + (int64_t)l * a.n_in + s];
            if (a.sgn && a.sgn[k * a.n_out + j] < 0 && v) v = (int64_t)a.q[l] - v;
        }
        out[i] = v;
    }
}
void ab_launch_tunnel_gather(const ab_plan *ps, int64_t *out, const int64_t *in, int64_t in_bstride, int64_t n_in, const int32_t *tab,
                             const int8_t *sgn, int64_t batch, int64_t K, int32_t nlimbs, cudaStream_t st) {
    if (batch == 0) return;
    TgArgs a{};
    a.nlimbs = nlimbs; a.n_out = ps->n; a.n_in = n_in; a.K = K; a.batch = batch; a.in_bstride = in_bstride; a.tab = tab; a.sgn = sgn;
    for (int l = 0; l < nlimbs; ++l) a.q[l] = ps->limb[l].q;
    k_tunnel_gather<<<grid_for(batch * K * nlimbs * ps->n, 256, ps->sm_count), 256, 0, st>>>(a, out, in);
    ab_count_launch();
    ab_check_launch("tunnel_gather");
}
struct ArArgs {
    uint32_t q[AB_MAX_LIMBS];
    int64_t n, per, batch, ob, ib;
};
__global__ void __launch_bounds__(256) k_add_rows(const ArArgs a, int64_t *__restrict__ out, const int64_t *__restrict__ in) {
    const int64_t total = a.batch * a.per;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / a.per, r = i % a.per;
        const uint32_t q = a.q[r / a.n];
        out[b * a.ob + r] = (int64_t)ab_addmod((uint32_t)out[b * a.ob + r], (uint32_t)in[b * a.ib + r], q);
    }
}
void ab_launch_add_rows(const ab_plan *p, int64_t *out, int64_t ob, const int64_t *in, int64_t ib, int64_t batch, int32_t nlimbs, cudaStream_t st) {
    if (batch == 0) return;
    ArArgs a{};
    a.n = p->n; a.per = (int64_t)nlimbs * p->n; a.batch = batch; a.ob = ob; a.ib = ib;
    for (int l = 0; l < nlimbs; ++l) a.q[l] = p->limb[l].q;
    k_add_rows<<<grid_for(batch * a.per, 256, p->sm_count), 256, 0, st>>>(a, out, in);
    ab_count_launch();
    ab_check_launch("add_rows");
}
