// ab_she.cpp -- ciphertext-level operations of Lol's SymmSHE as sequences of batched kernels.
//
// The reference reaches them at Crypto/Alchemy/Interpreter/Eval.hs:65-67 ((*) on CT), :130
// (modSwitch), :133 (keySwitchQuadCirc), :134 (tunnel); PT2CT.hs:172-177 and :224-229 fix the order
// in which they are chained.  Upstream performs them in Haskell on top of one-element Tensor calls
// [RECALL]; here each is a handful of launches over the whole batch, with the few intermediates that
// leave the SMs in a caller-supplied scratch buffer (`*_work_bytes`; very large batches are processed in
// chunks of that size).  All polynomials are in the CRT basis, PLANAR layout: c[b][i][l][j].
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "ab_internal.h"
#include "ab_kernels.h"

// scratch per chunk: every launch of a chunk should be many waves of CTAs deep (one CTA of the fused inner product runs for
// milliseconds, so a chunk of less than a few waves leaves SMs idle in its last wave)
static const int64_t kChunkBytes = 2048ll << 20;
static const int64_t kMacChunkBytes = 4096ll << 20;   // the fused paths keep only O(1) polynomials per ciphertext in scratch

static int64_t chunk_of(int64_t batch, int64_t bytes_per_item) {
    int64_t c = kChunkBytes / std::max<int64_t>(1, bytes_per_item);
    if (c < 1) c = 1;
    return std::min(batch, c);
}

// digit table of the product-ring gadget over the first `nlimbs` moduli (Lol: decompose (a,b) = decompose a ++ decompose b)
static void fill_digits(const ab_plan *p, int32_t nlimbs, int32_t base, AbMxFused &fx, bool single_limb_src, int32_t src_limbs = -1) {
    // src_limbs < nlimbs: the source is known to be ZERO in the limbs from src_limbs on (a ciphertext that was just scaled up to a
    // larger modulus): their gadget digits are zero, contribute nothing, and are left out (a prefix of the digit list remains)
    int32_t nd = 0;
    if (src_limbs >= 0 && src_limbs < nlimbs) nlimbs = src_limbs;
    for (int l = 0; l < nlimbs; ++l) {
        const int32_t k = ab_gadget_digits(p->q[l], base);
        uint64_t pw = 1, top = 1;                       // top = b^(k-1)
        for (int i = 0; i + 1 < k; ++i) top *= (uint64_t)base;
        const uint64_t H = base ? (uint64_t)(base / 2) * (top - 1) / (uint64_t)(base - 1) : 0;
        for (int i = 0; i < k; ++i, ++nd) {
            AB_REQUIRE(nd < AB_MX_MAX_DIGITS, AB_ERR_UNSUPPORTED, "too many gadget digits");
            AB_REQUIRE(pw < (1ull << 32) && H < (1ull << 32), AB_ERR_UNSUPPORTED, "gadget base too large");
            fx.dig[nd].q_src = (uint32_t)p->q[l];
            fx.dig[nd].limb = (int16_t)(single_limb_src ? 0 : l);
            fx.dig[nd].pw = (uint32_t)pw;
            fx.dig[nd].H = (uint32_t)H;
            fx.dig[nd].last = (int16_t)(i == k - 1);
            pw *= (uint64_t)(base ? base : 1);
        }
    }
    fx.nd = nd;
    fx.gbase = base;
}
static bool fused_ok(const ab_plan *p, int32_t nlimbs, const void *a, const void *b) {
    static const bool off = getenv("AB_NO_FUSE") && atoi(getenv("AB_NO_FUSE")) != 0;      // A/B measurements only: plain launch sequences
    if (off) return false;
    return !ab_pow2_supported(p) && ab_mixed_supported(p, AbView{(int64_t *)a, 1, 0, nlimbs, AB_LAYOUT_PLANAR}) && (((uintptr_t)a | (uintptr_t)b) & 15) == 0;
}

static bool mac_ok(const ab_plan *p, int32_t nlimbs, const void *a, const void *b) {
    static const bool off = getenv("AB_NO_MAC") && atoi(getenv("AB_NO_MAC")) != 0;      // A/B measurements only
    return !off && fused_ok(p, nlimbs, a, b) && ab_mixed_mac_supported(p, nlimbs);
}

extern "C" int ab_ct_mul_dev(const ab_plan *p, int64_t *out3, const int64_t *a2, const int64_t *b2, int64_t batch, int32_t nlimbs, void *stream) {
    return ab_guard([&] {
        AB_REQUIRE(p && ((out3 && a2 && b2) || batch == 0), AB_ERR_ARG, "null argument");
        AB_REQUIRE(p->has_crt, AB_ERR_NO_ROOT, "plan has no CRT tables");
        AB_REQUIRE(batch >= 0 && nlimbs >= 1 && nlimbs <= p->L, AB_ERR_ARG, "bad batch or limb count");
        ab_launch_ct_mul(p, out3, a2, b2, batch, nlimbs, (cudaStream_t)stream);
    });
}

// ------------------------------------------------------------------------------------ modSwitch
// rows = batch * npoly polynomials; per row: scratch for the dropped limb (n), the lifted limbs
// ((from-1) n) and two ping-pong copies of the shrinking polynomial ((from-1) n each).
static int64_t ms_row_words(const ab_plan *p, int32_t from, int32_t to) {
    if (to >= from) return 0;
    return p->n * (1 + 3 * (int64_t)(from - 1));
}
extern "C" int64_t ab_mod_switch_work_bytes(const ab_plan *p, int64_t batch, int32_t npoly, int32_t from, int32_t to) {
    if (!p || batch < 0 || npoly < 1) return -1;
    const int64_t rw = ms_row_words(p, from, to);
    return chunk_of(batch * npoly, rw * 8) * rw * 8;
}
// Which basis each polynomial of a ciphertext is rescaled in: bit i of dec_mask set = polynomial i goes through the DECODING basis
// (L^-1 after the inverse transform of the dropped limbs, L before the forward transform of the correction), clear = powerful basis.
// Rows are polynomials: row r of the call is polynomial (r + phase) % period of its ciphertext.  Rings without odd primes have L = id.
struct MsBasis {
    int32_t period = 1;
    uint32_t mask = 0;
    bool any(const ab_plan *p) const {
        if ((mask & ((period >= 32 ? 0u : (1u << period)) - 1u)) == 0) return false;
        for (const AbAxis &ax : p->axes) if (ax.p != 2) return true;
        return false;
    }
    void apply(AbMxFused &f, bool inverse, int64_t row0) const {
        if (inverse) f.post_linv = 1; else f.pre_l = 1;
        f.l_period = period; f.l_mask = (int32_t)mask; f.l_phase = (int32_t)(row0 % period);
    }
};

// `rows` polynomials; row r is read at in + r*in_stride (from limbs) and written at out + r*out_stride (to limbs)
static void mod_switch_rows(const ab_plan *p, int64_t *out, int64_t out_stride, const int64_t *in, int64_t in_stride, int64_t rows, int32_t from, int32_t to,
                            int64_t *work, cudaStream_t st, const MsBasis &basis, bool can_fuse) {
    const int64_t n = p->n;
    if (to >= from) {
        if (to == from) ab_launch_copy_rows(out, out_stride, in, in_stride, from * n, rows, st);
        else ab_launch_rescale_up(p, out, out_stride, in, in_stride, rows, from, to, st);
        return;
    }
    const bool dec = basis.any(p);
    AB_REQUIRE(!dec || can_fuse || basis.period == 1, AB_ERR_UNSUPPORTED, "mixed rescale bases need the per-polynomial route");
    const int64_t rw = ms_row_words(p, from, to);
    const int64_t chunk = chunk_of(rows, rw * 8);
    for (int64_t r0 = 0; r0 < rows; r0 += chunk) {
        const int64_t R = std::min(chunk, rows - r0);
        int64_t *dropped = work, *lifted = dropped + R * n, *ping = lifted + R * (from - 1) * n, *pong = ping + R * (from - 1) * n;
        const int64_t *cur = in + r0 * in_stride;
        int64_t cur_stride = in_stride;
        for (int32_t L = from; L > to; --L) {
            const int32_t d = L - 1;
            static const bool no_multi = getenv("AB_NO_MULTIDROP") && atoi(getenv("AB_NO_MULTIDROP")) != 0;      // A/B measurements only
            if (can_fuse && !no_multi && L - to >= 2) {
                // drop k = min(L - to, 4) limbs in ONE rescale: crtInv of the k dropped limbs in one launch, then one kernel per
                // batch that rebuilds, per coefficient, the centred digits of the sequential drops (Garner), reduces
                // x = r_0 + q_0 (r_1 + ...) into each kept modulus, transforms it and applies (in - x) (q_0 q_1 ...)^-1:
                // k inverse + (L - k) forward transforms per polynomial instead of k + (L-1) + (L-2) + ...
                const int32_t k = std::min<int32_t>(L - to, AB_MX_MAX_DROP), keep = L - k;
                AbMxFused fi{};
                fi.load_mode = AB_MX_LOAD_SRC; fi.src = cur; fi.src_ustride = cur_stride; fi.src_limb0 = keep;
                if (dec) basis.apply(fi, true, r0);
                ab_launch_crt_mixed_fused(p, AbView{dropped, R, keep, k, AB_LAYOUT_PLANAR}, true, &fi, st);    // [R][k][n], ascending limbs
                AbMxFused fx{};
                fx.load_mode = AB_MX_LOAD_GARNER; fx.store_mode = AB_MX_STORE_SUBSCALE;
                if (dec) basis.apply(fx, false, r0);
                // drop order: i = 0 is limb L-1; the kernel reads src + i*n, so point it at the last limb and step backwards
                fx.src = dropped + (int64_t)(k - 1) * n; fx.src_ustride = (int64_t)k * n; fx.gar_k = k;
                for (int i = 0; i < k; ++i) {
                    const uint64_t qi = (uint64_t)p->q[L - 1 - i];
                    fx.gar_q[i] = (uint32_t)qi;
                    for (int t2 = i + 1; t2 < k; ++t2) {
                        const uint64_t qt = (uint64_t)p->q[L - 1 - t2];
                        const uint64_t inv = abh::invmod(qi % qt, qt);
                        fx.gar_inv[i][t2] = (uint32_t)inv; fx.gar_inv_s[i][t2] = abh::shoup(inv, qt);
                    }
                }
                fx.sub_in = cur; fx.sub_ustride = cur_stride;
                for (int kk = 0; kk < keep; ++kk) {
                    const uint64_t qk = (uint64_t)p->q[kk];
                    uint64_t prod = 1;
                    for (int i = 0; i < k; ++i) prod = abh::mulmod(prod, (uint64_t)p->q[L - 1 - i] % qk, qk);
                    const uint64_t inv = abh::invmod(prod, qk);
                    fx.scale[kk] = (uint32_t)inv; fx.scale_s[kk] = abh::shoup(inv, qk);
                }
                int64_t *dst_buf;
                if (keep == to) { dst_buf = out + r0 * out_stride; fx.dst_ustride = out_stride; }
                else dst_buf = (cur == ping) ? pong : ping;
                ab_launch_crt_mixed_fused(p, AbView{dst_buf, R, 0, keep, AB_LAYOUT_PLANAR}, false, &fx, st);
                cur = dst_buf;
                cur_stride = keep == to ? out_stride : (int64_t)keep * n;
                L = keep + 1;          // the loop's --L lands on `keep`
                continue;
            }
            if (can_fuse) {
                // two launches per dropped limb: crtInv of that limb (out of place), then one kernel that lifts it,
                // reduces it into each kept modulus, transforms it and applies (in - x) * q_d^-1 on the way out
                AbMxFused fi{};
                fi.load_mode = AB_MX_LOAD_SRC; fi.src = cur; fi.src_ustride = cur_stride; fi.src_limb0 = d;   // view limb0 = d selects the modulus
                if (dec) basis.apply(fi, true, r0);
                ab_launch_crt_mixed_fused(p, AbView{dropped, R, d, 1, AB_LAYOUT_PLANAR}, true, &fi, st);
                AbMxFused fx{};
                fx.load_mode = AB_MX_LOAD_DIGITS; fx.store_mode = AB_MX_STORE_SUBSCALE;
                if (dec) basis.apply(fx, false, r0);
                fx.nd = 1; fx.gbase = 0; fx.src = dropped; fx.src_ustride = n; fx.n_src = (int32_t)n; fx.gK = 1;
                fx.dig[0].q_src = (uint32_t)p->q[d]; fx.dig[0].limb = 0; fx.dig[0].pw = 1; fx.dig[0].H = 0; fx.dig[0].last = 1;
                fx.sub_in = cur; fx.sub_ustride = cur_stride;
                for (int k = 0; k < d; ++k) {
                    const uint64_t inv = abh::invmod((uint64_t)p->q[d] % (uint64_t)p->q[k], (uint64_t)p->q[k]);
                    fx.scale[k] = (uint32_t)inv; fx.scale_s[k] = abh::shoup(inv, (uint64_t)p->q[k]);
                }
                int64_t *dst_buf;
                if (d == to) { dst_buf = out + r0 * out_stride; fx.dst_ustride = out_stride; }
                else dst_buf = (cur == ping) ? pong : ping;
                ab_launch_crt_mixed_fused(p, AbView{dst_buf, R, 0, d, AB_LAYOUT_PLANAR}, false, &fx, st);
                cur = dst_buf;
                cur_stride = d == to ? out_stride : (int64_t)d * n;
                continue;
            }
            // launch sequence (power-of-two rings, rings outside the shared-memory kernel): here every row shares one basis
            ab_launch_copy_rows(dropped, n, cur + (int64_t)d * n, cur_stride, n, R, st);
            ab_crt_rows(p, dropped, R, d, true, st);
            if (dec) ab_launch_prime_op(p, AB_PR_LINV, AbView{dropped, R, d, 1, AB_LAYOUT_PLANAR}, st);
            int32_t dst[AB_MAX_LIMBS];
            for (int k = 0; k < d; ++k) dst[k] = k;
            ab_launch_lift_reduce(p, lifted, (int64_t)d * n, dropped, n, R, d, dst, d, st);
            if (dec) ab_launch_prime_op(p, AB_PR_L, AbView{lifted, R, 0, d, AB_LAYOUT_PLANAR}, st);
            ab_crt_dispatch(p, AbView{lifted, R, 0, d, AB_LAYOUT_PLANAR}, false, st);
            int64_t scal[AB_MAX_LIMBS];
            for (int k = 0; k < d; ++k) scal[k] = (int64_t)abh::invmod((uint64_t)p->q[d] % (uint64_t)p->q[k], (uint64_t)p->q[k]);
            int64_t *dst_buf;
            int64_t dst_stride;
            if (d == to) { dst_buf = out + r0 * out_stride; dst_stride = out_stride; }
            else { dst_buf = (cur == ping) ? pong : ping; dst_stride = (int64_t)d * n; }
            ab_launch_sub_scale(p, dst_buf, dst_stride, cur, cur_stride, lifted, (int64_t)d * n, R, d, scal, st);
            cur = dst_buf;
            cur_stride = dst_stride;
        }
    }
}
static void mod_switch_polys(const ab_plan *p, int64_t *out, const int64_t *in, int64_t batch, int32_t npoly, int32_t from, int32_t to, uint32_t dec_mask,
                             int64_t *work, cudaStream_t st) {
    const int64_t n = p->n;
    MsBasis basis;
    basis.period = npoly; basis.mask = dec_mask;
    const bool can_fuse = to < from && fused_ok(p, from, in, work) && ((uintptr_t)out & 15) == 0;
    if (to >= from || !basis.any(p) || can_fuse) {
        mod_switch_rows(p, out, (int64_t)to * n, in, (int64_t)from * n, batch * npoly, from, to, work, st, basis, can_fuse);
        return;
    }
    // launch-sequence route with polynomials in different bases: one pass per polynomial index, rows strided by the ciphertext
    for (int32_t i = 0; i < npoly; ++i) {
        MsBasis one;
        one.period = 1; one.mask = (dec_mask >> i) & 1u;
        mod_switch_rows(p, out + (int64_t)i * to * n, (int64_t)npoly * to * n, in + (int64_t)i * from * n, (int64_t)npoly * from * n, batch, from, to, work, st, one, false);
    }
}
extern "C" int ab_mod_switch_dev(const ab_plan *p, int64_t *out, const int64_t *in, int64_t batch, int32_t npoly, int32_t from, int32_t to,
                                 uint32_t dec_mask, void *work, void *stream) {
    return ab_guard([&] {
        AB_REQUIRE(p && ((out && in) || batch == 0), AB_ERR_ARG, "null argument");
        AB_REQUIRE(p->has_crt, AB_ERR_NO_ROOT, "plan has no CRT tables");
        AB_REQUIRE(batch >= 0 && npoly >= 1 && npoly <= 31 && from >= 1 && to >= 1 && from <= p->L && to <= p->L, AB_ERR_ARG, "bad limb counts");
        AB_REQUIRE(to >= from || work != nullptr, AB_ERR_ARG, "null scratch buffer");
        mod_switch_polys(p, out, in, batch, npoly, from, to, dec_mask, (int64_t *)work, (cudaStream_t)stream);
    });
}

// ------------------------------------------------------------------------------------ keySwitchQuadCirc
static int64_t ks_item_words(const ab_plan *p, int32_t nlimbs, int32_t ndig) { return p->n * nlimbs * (1 + (int64_t)ndig); }
static int64_t ks_mac_chunk(const ab_plan *p, int64_t batch, int32_t nlimbs) {
    return std::min<int64_t>(batch, std::max<int64_t>(1, kMacChunkBytes / (3 * (int64_t)nlimbs * p->n * 8)));
}
extern "C" int64_t ab_key_switch_work_bytes(const ab_plan *p, int64_t batch, int32_t nlimbs, int32_t base) {
    const int32_t nd = ab_gadget_len(p, nlimbs, base);
    if (nd < 0 || batch < 0) return -1;
    const int64_t w = ks_item_words(p, nlimbs, nd);
    const int64_t generic = chunk_of(batch, w * 8) * w * 8;
    const int64_t fused = ks_mac_chunk(p, batch, nlimbs) * 3 * nlimbs * p->n * 8;          // c2, running sums, centred source
    return std::max(generic, fused) + (int64_t)nd * 2 * nlimbs * p->n * 8;     // + the prepared hint
}
static int key_switch_quad(const ab_plan *p, int64_t *out2, const int64_t *in3, const int64_t *hint, int64_t batch, int32_t nlimbs,
                           int32_t base, void *work, void *stream, int32_t src_limbs);
extern "C" int ab_key_switch_quad_dev(const ab_plan *p, int64_t *out2, const int64_t *in3, const int64_t *hint, int64_t batch, int32_t nlimbs,
                                      int32_t base, void *work, void *stream) {
    return key_switch_quad(p, out2, in3, hint, batch, nlimbs, base, work, stream, nlimbs);
}
// src_limbs: c2 is known to vanish in the limbs from src_limbs on (see fill_digits); nlimbs = no such knowledge
static int key_switch_quad(const ab_plan *p, int64_t *out2, const int64_t *in3, const int64_t *hint, int64_t batch, int32_t nlimbs,
                           int32_t base, void *work, void *stream, int32_t src_limbs) {
    return ab_guard([&] {
        AB_REQUIRE(p && ((out2 && in3 && hint && work) || batch == 0), AB_ERR_ARG, "null argument");
        AB_REQUIRE(p->has_crt, AB_ERR_NO_ROOT, "plan has no CRT tables");
        const int32_t nd = ab_gadget_len(p, nlimbs, base);
        AB_REQUIRE(nd > 0 && batch >= 0, AB_ERR_ARG, "bad limb count or gadget base");
        cudaStream_t st = (cudaStream_t)stream;
        const int64_t n = p->n, per = (int64_t)nlimbs * n;
        const int64_t chunk = chunk_of(batch, ks_item_words(p, nlimbs, nd) * 8);
        int64_t *hm = (int64_t *)work;                         // prepared hint: [nd][2][L][n] uint2
        int64_t *work2 = hm + (int64_t)nd * 2 * per;
        if (batch > 0 && mac_ok(p, nlimbs, in3, work)) {
            // two launches per chunk: crtInv of c2 straight out of the input, then the fused gadget inner product
            // (digit extraction, forward CRT and multiply-accumulate on chip; ab_crt_mixed.cu k_crt_mixed_mac)
            ab_launch_hint_prep2(p, hm, hint, nd, nlimbs, st);
            const int64_t chunk2 = ks_mac_chunk(p, batch, nlimbs);     // scratch: c2, running sums, centred source
            for (int64_t b0 = 0; b0 < batch; b0 += chunk2) {
                const int64_t B = std::min(chunk2, batch - b0);
                int64_t *c2 = work2, *accs = c2 + B * per, *tb = accs + B * per;
                const int64_t *in = in3 + b0 * 3 * per;
                AbMxFused fi{};
                fi.load_mode = AB_MX_LOAD_SRC; fi.src = in + 2 * per; fi.src_ustride = 3 * per; fi.src_limb0 = 0;
                fi.dst_ustride = per;                                  // only the limbs that can be non-zero are brought to the powerful basis
                ab_launch_crt_mixed_fused(p, AbView{c2, B, 0, src_limbs, AB_LAYOUT_PLANAR}, true, &fi, st);
                AbMxFused fx{};
                fx.load_mode = AB_MX_LOAD_DIGITS;
                fill_digits(p, nlimbs, base, fx, false, src_limbs);
                fx.src = c2; fx.src_ustride = per; fx.n_src = (int32_t)n; fx.gK = 1;
                if (base == 0) { fx.triv_src = in + 2 * per; fx.triv_ustride = 3 * per; }
                AbMxMac mac{};
                mac.hm = hm; mac.K = 1; mac.add = in; mac.add_ustride = 3 * per; mac.acc = accs; mac.tbuf = tb;
                mac.out = out2 + b0 * 2 * per; mac.out_ustride = 2 * per;
                ab_launch_crt_mixed_mac(p, B, nlimbs, fx, mac, st);
            }
            return;
        }
        if (batch > 0) ab_launch_hint_prep(p, hm, hint, (int64_t)nd * 2, nlimbs, st);
        for (int64_t b0 = 0; b0 < batch; b0 += chunk) {
            const int64_t B = std::min(chunk, batch - b0);
            int64_t *c2 = work2, *dig = c2 + B * per;
            const int64_t *in = in3 + b0 * 3 * per;
            if (fused_ok(p, nlimbs, in, work)) {
                // crtInv of c2 straight out of the input; then ONE launch that takes each gadget digit in the load
                // phase of the forward CRT (TrivGad digits of limb l mod q_l are the CRT-basis input itself)
                AbMxFused fi{};
                fi.load_mode = AB_MX_LOAD_SRC; fi.src = in + 2 * per; fi.src_ustride = 3 * per; fi.src_limb0 = 0;
                ab_launch_crt_mixed_fused(p, AbView{c2, B, 0, nlimbs, AB_LAYOUT_PLANAR}, true, &fi, st);
                AbMxFused fx{};
                fx.load_mode = AB_MX_LOAD_DIGITS;
                fill_digits(p, nlimbs, base, fx, false);
                fx.src = c2; fx.src_ustride = per; fx.n_src = (int32_t)n; fx.gK = 1;
                if (base == 0) { fx.triv_src = in + 2 * per; fx.triv_ustride = 3 * per; }
                ab_launch_crt_mixed_fused(p, AbView{dig, B * nd, 0, nlimbs, AB_LAYOUT_PLANAR}, false, &fx, st);
                ab_launch_hint_ip_m(p, out2 + b0 * 2 * per, 2 * per, dig, (int64_t)nd * per, hm, in, 3 * per, B, nlimbs, nd, st);
                continue;
            }
            // c2 -> powerful basis -> gadget digits reduced into every limb -> back to CRT
            ab_launch_copy_rows(c2, per, in + 2 * per, 3 * per, per, B, st);
            ab_crt_dispatch(p, AbView{c2, B, 0, nlimbs, AB_LAYOUT_PLANAR}, true, st);
            ab_launch_decompose(p, dig, c2, B, nlimbs, nlimbs, base, st);
            ab_crt_dispatch(p, AbView{dig, B * nd, 0, nlimbs, AB_LAYOUT_PLANAR}, false, st);
            // (c0, c1) + sum_d dig_d * hint_d
            ab_launch_hint_ip_m(p, out2 + b0 * 2 * per, 2 * per, dig, (int64_t)nd * per, hm, in, 3 * per, B, nlimbs, nd, st);
        }
    });
}

// ------------------------------------------------------------------------------------ tunnel
struct ab_tunnel {
    const ab_plan *r = nullptr, *e = nullptr, *s = nullptr;
    ab_ext *er = nullptr, *es = nullptr;
    int64_t K = 0;                 // dimension of O_r' over O_e'
    int32_t *d_pow = nullptr;      // [K][n_s] index into [n_r] of the powerful-basis source, -1 = zero
    int32_t *d_dec = nullptr;      // same for the decoding basis
    int8_t *d_dec_sgn = nullptr;   // [K][n_s]
    // sub-ring route for the c1 part: CRT_s'(embed x) = embedCRT(CRT_e'(x)), so the gadget digit transforms run in O_e' (a plan
    // with CRT tables for e' over the moduli of s', root omega_s'^(s'/e')) and are replicated through es_crt->d_base_crt
    ab_plan *ecrt = nullptr;
    ab_ext *es_crt = nullptr;
    bool rep4 = false;             // es_crt->d_base_crt maps every aligned group of four rows of O_s' to four consecutive, aligned rows of O_e'
};

extern "C" int ab_tunnel_create(ab_tunnel **out, const ab_plan *plan_r, const ab_plan *plan_e, const ab_plan *plan_s) {
    ab_tunnel *t = nullptr;
    int rc = ab_guard([&] {
        AB_REQUIRE(out && plan_r && plan_e && plan_s, AB_ERR_ARG, "null argument");
        AB_REQUIRE(plan_r->has_crt && plan_s->has_crt, AB_ERR_NO_ROOT, "r' and s' plans need CRT tables");
        t = new ab_tunnel();
        t->r = plan_r; t->e = plan_e; t->s = plan_s;
        int rc2 = ab_ext_create(&t->er, plan_e, plan_r);
        if (rc2 != AB_OK) throw ab_error{rc2, ab_last_error()};
        rc2 = ab_ext_create(&t->es, plan_e, plan_s);
        if (rc2 != AB_OK) throw ab_error{rc2, ab_last_error()};
        t->K = t->er->rel;
        const int64_t ne = plan_e->n, ns = plan_s->n, K = t->K;
        // compose coeffs (r' -> K x e') with embed (e' -> s') into one gather table per basis
        std::vector<int32_t> coeff(K * ne), pow_to_lo(ns), edec_idx(ns);
        std::vector<int8_t> edec_sgn(ns);
        AB_CHECK_CUDA(cudaMemcpy(coeff.data(), t->er->d_coeff, coeff.size() * 4, cudaMemcpyDeviceToHost));
        AB_CHECK_CUDA(cudaMemcpy(pow_to_lo.data(), t->es->d_pow_to_lo, ns * 4, cudaMemcpyDeviceToHost));
        AB_CHECK_CUDA(cudaMemcpy(edec_idx.data(), t->es->d_edec_idx, ns * 4, cudaMemcpyDeviceToHost));
        AB_CHECK_CUDA(cudaMemcpy(edec_sgn.data(), t->es->d_edec_sgn, ns, cudaMemcpyDeviceToHost));
        std::vector<int32_t> tp(K * ns), td(K * ns);
        std::vector<int8_t> ts(K * ns);
        for (int64_t k = 0; k < K; ++k)
            for (int64_t j = 0; j < ns; ++j) {
                tp[k * ns + j] = pow_to_lo[j] < 0 ? -1 : coeff[k * ne + pow_to_lo[j]];
                td[k * ns + j] = edec_idx[j] < 0 ? -1 : coeff[k * ne + edec_idx[j]];
                ts[k * ns + j] = edec_sgn[j];
            }
        AB_CHECK_CUDA(cudaMalloc((void **)&t->d_pow, tp.size() * 4));
        AB_CHECK_CUDA(cudaMalloc((void **)&t->d_dec, td.size() * 4));
        AB_CHECK_CUDA(cudaMalloc((void **)&t->d_dec_sgn, ts.size()));
        AB_CHECK_CUDA(cudaMemcpy(t->d_pow, tp.data(), tp.size() * 4, cudaMemcpyHostToDevice));
        AB_CHECK_CUDA(cudaMemcpy(t->d_dec, td.data(), td.size() * 4, cudaMemcpyHostToDevice));
        AB_CHECK_CUDA(cudaMemcpy(t->d_dec_sgn, ts.data(), ts.size(), cudaMemcpyHostToDevice));
        // CRT plan of e' over the moduli of s' (e' | s' | q - 1), same root family; optional: without it the digits run in O_s'
        if (plan_e->m > 2 && K >= 1 && ns > ne) {
            std::vector<ab_prime_exp> pe;
            for (const AbAxis &ax : plan_e->axes) pe.push_back(ab_prime_exp{ax.p, ax.e});
            std::vector<int64_t> roots;
            for (int l = 0; l < plan_s->L; ++l)
                roots.push_back((int64_t)abh::powmod((uint64_t)plan_s->root[l], (uint64_t)(plan_s->m / plan_e->m), (uint64_t)plan_s->q[l]));
            if (ab_plan_create(&t->ecrt, pe.data(), (int32_t)pe.size(), plan_s->q.data(), plan_s->L, roots.data()) != AB_OK) t->ecrt = nullptr;
            else if (ab_ext_create(&t->es_crt, t->ecrt, plan_s) != AB_OK) { ab_plan_destroy(t->ecrt); t->ecrt = nullptr; t->es_crt = nullptr; }
            if (t->es_crt && (ns & 3) == 0) {
                std::vector<int32_t> base(ns);
                AB_CHECK_CUDA(cudaMemcpy(base.data(), t->es_crt->d_base_crt, ns * 4, cudaMemcpyDeviceToHost));
                bool ok = true;
                for (int64_t j = 0; j < ns && ok; j += 4)
                    ok = (base[j] & 3) == 0 && base[j + 1] == base[j] + 1 && base[j + 2] == base[j] + 2 && base[j + 3] == base[j] + 3;
                static const bool off = getenv("AB_NO_REP4") && atoi(getenv("AB_NO_REP4")) != 0;      // A/B measurements only
                t->rep4 = ok && !off;
            }
        }
        *out = t;
    });
    if (rc != AB_OK && t) ab_tunnel_destroy(t);
    return rc;
}
extern "C" void ab_tunnel_destroy(ab_tunnel *t) {
    if (!t) return;
    ab_ext_destroy(t->er);
    ab_ext_destroy(t->es);
    if (t->es_crt) ab_ext_destroy(t->es_crt);
    if (t->ecrt) ab_plan_destroy(t->ecrt);
    cudaFree(t->d_pow);
    cudaFree(t->d_dec);
    cudaFree(t->d_dec_sgn);
    delete t;
}

// scratch per ciphertext: crtInv'd input 2 L n_r; decoding-basis copy of c0 L n_r; c0 parts and c1 parts
// K L n_s each; digits K nd L n_s; the evalLin result L n_s
static int64_t tn_item_words(const ab_tunnel *t, int32_t nlimbs, int32_t nd) {
    return 3 * (int64_t)nlimbs * t->r->n + (int64_t)nlimbs * t->s->n * (t->K * (2 + nd) + 1);
}
// fused path: crtInv'd input 2 L n_r, decoding-basis c0 L n_r, c0 parts K L n_s, evalLin result, running sums, centred source
static int64_t tn_mac_words(const ab_tunnel *t, int32_t nlimbs, int32_t parts) {
    return 3 * (int64_t)nlimbs * t->r->n + (t->K + 1 + 2 * (int64_t)parts) * (int64_t)nlimbs * t->s->n;
}
static int64_t tn_mac_chunk(const ab_tunnel *t, int64_t batch, int32_t nlimbs, int32_t parts) {
    return std::min<int64_t>(batch, std::max<int64_t>(1, kMacChunkBytes / (tn_mac_words(t, nlimbs, parts) * 8)));
}
extern "C" int64_t ab_tunnel_work_bytes(const ab_tunnel *t, int64_t batch, int32_t nlimbs, int32_t base) {
    if (!t || batch < 0) return -1;
    const int32_t nd = ab_gadget_len(t->s, nlimbs, base);
    if (nd < 0) return -1;
    const int64_t w = tn_item_words(t, nlimbs, nd);
    const int64_t generic = chunk_of(batch, w * 8) * w * 8;
    const int32_t parts = ab_mixed_mac_parts(t->s, batch, nlimbs, (int32_t)(t->K * nd));
    const int64_t fused = tn_mac_chunk(t, batch, nlimbs, parts) * tn_mac_words(t, nlimbs, parts) * 8;
    return std::max(generic, fused) + t->K * (int64_t)nd * 2 * nlimbs * t->s->n * 8;      // + the prepared hint
}

extern "C" int ab_tunnel_dev(const ab_tunnel *t, int64_t *out2_s, const int64_t *in2_r, const int64_t *flin, const int64_t *hint, int64_t batch,
                             int32_t nlimbs, int32_t base, void *work, void *stream) {
    return ab_guard([&] {
        AB_REQUIRE(t && ((out2_s && in2_r && flin && hint && work) || batch == 0), AB_ERR_ARG, "null argument");
        const ab_plan *pr = t->r, *ps = t->s;
        AB_REQUIRE(batch >= 0 && nlimbs >= 1 && nlimbs <= pr->L && nlimbs <= ps->L, AB_ERR_ARG, "bad limb count");
        const int32_t nd = ab_gadget_len(ps, nlimbs, base);
        AB_REQUIRE(nd > 0, AB_ERR_ARG, "bad gadget base");
        for (int l = 0; l < nlimbs; ++l) AB_REQUIRE(pr->q[l] == ps->q[l], AB_ERR_ARG, "r' and s' plans must share their moduli");
        cudaStream_t st = (cudaStream_t)stream;
        const int64_t nr = pr->n, ns = ps->n, K = t->K, ne = t->e->n;
        const int64_t per_r = (int64_t)nlimbs * nr, per_s = (int64_t)nlimbs * ns;
        const int64_t chunk = chunk_of(batch, tn_item_words(t, nlimbs, nd) * 8);
        int64_t *hm = (int64_t *)work;               // prepared hint: [K][nd][2][L][n_s] uint2
        int64_t *work2 = hm + K * (int64_t)nd * 2 * per_s;
        if (batch > 0 && mac_ok(ps, nlimbs, out2_s, work) && ((per_r * 8) % 16) == 0) {
            ab_launch_hint_prep2(ps, hm, hint, K * (int64_t)nd, nlimbs, st);
            const int32_t parts = ab_mixed_mac_parts(ps, batch, nlimbs, (int32_t)(K * nd));
            const int64_t chunk2 = tn_mac_chunk(t, batch, nlimbs, parts);
            for (int64_t b0 = 0; b0 < batch; b0 += chunk2) {
                const int64_t B = std::min(chunk2, batch - b0);
                int64_t *cin = work2;                    // [B][2][L][n_r]
                int64_t *c0d = cin + B * 2 * per_r;      // [B][L][n_r]
                int64_t *p0 = c0d + B * per_r;           // [B][K][L][n_s]
                int64_t *lin = p0 + B * K * per_s;       // [B][L][n_s]
                int64_t *accs = lin + B * per_s;         // [B][parts][L][n_s] running sums of the fused inner product
                int64_t *tb = accs + B * parts * per_s;  // [B][parts][L][n_s] centred source + H of the current (k, source limb)
                const int64_t *inb = in2_r + b0 * 2 * per_r;
                static const bool no_sub = getenv("AB_NO_SUBRING") && atoi(getenv("AB_NO_SUBRING")) != 0;      // A/B measurements only
                const bool sub_ok = !no_sub && t->ecrt && (ns & 3) == 0 && !ab_pow2_supported(t->ecrt) && ab_mixed_mac_supported(t->ecrt, nlimbs);
                bool fused_c0 = false;
                if (fused_ok(pr, nlimbs, inb, cin) && ((2 * per_r * 8) % 16) == 0) {
                    // c0 straight to the decoding basis (crtInv with L^-1 folded in), c1 to the powerful basis: two out-of-place
                    // launches, no copies, no separate L / L^-1 passes; L is folded into the load phase of the forward CRT of the parts
                    AbMxFused f0{};
                    f0.load_mode = AB_MX_LOAD_SRC; f0.src = inb; f0.src_ustride = 2 * per_r; f0.src_limb0 = 0; f0.post_linv = 1;
                    ab_launch_crt_mixed_fused(pr, AbView{c0d, B, 0, nlimbs, AB_LAYOUT_PLANAR}, true, &f0, st);
                    AbMxFused f1{};
                    f1.load_mode = AB_MX_LOAD_SRC; f1.src = inb + per_r; f1.src_ustride = 2 * per_r; f1.src_limb0 = 0;
                    // c1 lands in the second half of every cin row so that the digit source below keeps its layout
                    f1.dst_ustride = 2 * per_r;
                    ab_launch_crt_mixed_fused(pr, AbView{cin + per_r, B, 0, nlimbs, AB_LAYOUT_PLANAR}, true, &f1, st);
                    if (sub_ok) {
                        // the E'-coefficients of c0 stay in O_e': gathered in the load phase, decoding -> powerful (L) and CRT_e' in one
                        // launch; evalLin reads them through the embedCRT index (CRT_s'(embed x) = replicate(CRT_e'(x)))
                        AbMxFused fg{};
                        fg.load_mode = AB_MX_LOAD_GATHER; fg.src = c0d; fg.src_ustride = per_r; fg.n_src = (int32_t)nr;
                        fg.gather = t->er->d_coeff; fg.gK = (int32_t)K; fg.pre_l = 1;
                        ab_launch_crt_mixed_fused(t->ecrt, AbView{p0, B * K, 0, nlimbs, AB_LAYOUT_PLANAR}, false, &fg, st);
                    } else {
                        ab_launch_tunnel_gather(ps, p0, c0d, per_r, nr, t->d_dec, t->d_dec_sgn, B, K, nlimbs, st);
                        AbMxFused fl{};
                        fl.pre_l = 1;
                        ab_launch_crt_mixed_fused(ps, AbView{p0, B * K, 0, nlimbs, AB_LAYOUT_PLANAR}, false, &fl, st);
                    }
                    fused_c0 = true;
                } else {
                    ab_launch_copy_rows(cin, 2 * per_r, inb, 2 * per_r, 2 * per_r, B, st);
                    ab_crt_dispatch(pr, AbView{cin, B * 2, 0, nlimbs, AB_LAYOUT_PLANAR}, true, st);
                    ab_launch_copy_rows(c0d, per_r, cin, 2 * per_r, per_r, B, st);
                    ab_launch_prime_op(pr, AB_PR_LINV, AbView{c0d, B, 0, nlimbs, AB_LAYOUT_PLANAR}, st);
                    ab_launch_tunnel_gather(ps, p0, c0d, per_r, nr, t->d_dec, t->d_dec_sgn, B, K, nlimbs, st);
                    ab_launch_prime_op(ps, AB_PR_L, AbView{p0, B * K, 0, nlimbs, AB_LAYOUT_PLANAR}, st);
                    ab_crt_dispatch(ps, AbView{p0, B * K, 0, nlimbs, AB_LAYOUT_PLANAR}, false, st);
                }
                if (fused_c0 && sub_ok) ab_launch_lin_ip_rep(ps, lin, per_s, p0, K * (int64_t)nlimbs * ne, flin, t->es_crt->d_base_crt, ne, B, nlimbs, (int32_t)K, st);
                else ab_launch_lin_ip(ps, lin, per_s, p0, K * per_s, flin, B, nlimbs, (int32_t)K, st);
                // c1: gather (coeffs . embed), gadget digit, reduction, forward CRT of O_s' and the inner product with
                // the hint, all inside one kernel; the evalLin part of c0 is added on the way out
                AbMxFused fx{};
                fx.load_mode = AB_MX_LOAD_DIGITS;
                fill_digits(ps, nlimbs, base, fx, false);
                fx.src = cin + per_r; fx.src_ustride = 2 * per_r; fx.n_src = (int32_t)nr; fx.gK = (int32_t)K;
                AbMxMac mac{};
                mac.hm = hm; mac.K = (int32_t)K; mac.add0 = lin; mac.add0_ustride = per_s; mac.acc = accs; mac.tbuf = tb; mac.parts = parts;
                mac.out = out2_s + b0 * 2 * per_s; mac.out_ustride = 2 * per_s;
                if (sub_ok) {
                    // digits of embed(coeffs_k(c1)) = embed(digits of coeffs_k(c1)): transform them in O_e' and replicate (embedCRT)
                    fx.gather = t->er->d_coeff;
                    mac.rep_idx = t->es_crt->d_base_crt; mac.n_out = (int32_t)ns; mac.rep4 = t->rep4 ? 1 : 0;
                    ab_launch_crt_mixed_mac(t->ecrt, B, nlimbs, fx, mac, st);
                } else {
                    fx.gather = t->d_pow;
                    ab_launch_crt_mixed_mac(ps, B, nlimbs, fx, mac, st);
                }
            }
            return;
        }
        if (batch > 0) ab_launch_hint_prep(ps, hm, hint, K * (int64_t)nd * 2, nlimbs, st);
        for (int64_t b0 = 0; b0 < batch; b0 += chunk) {
            const int64_t B = std::min(chunk, batch - b0);
            int64_t *cin = work2;                    // [B][2][L][n_r]
            int64_t *c0d = cin + B * 2 * per_r;      // [B][L][n_r]
            int64_t *p0 = c0d + B * per_r;           // [B][K][L][n_s]
            int64_t *p1 = p0 + B * K * per_s;        // [B][K][L][n_s]
            int64_t *dig = p1 + B * K * per_s;       // [B][K][nd][L][n_s]
            int64_t *lin = dig + B * K * nd * per_s; // [B][L][n_s]
            int64_t *o = out2_s + b0 * 2 * per_s;
            // (c0, c1) to the powerful basis of O_r'
            ab_launch_copy_rows(cin, 2 * per_r, in2_r + b0 * 2 * per_r, 2 * per_r, 2 * per_r, B, st);
            ab_crt_dispatch(pr, AbView{cin, B * 2, 0, nlimbs, AB_LAYOUT_PLANAR}, true, st);
            // c0: decoding basis, split over E' (coeffsDec), embed each part into O_s' (embedDec), to CRT
            ab_launch_copy_rows(c0d, per_r, cin, 2 * per_r, per_r, B, st);
            ab_launch_prime_op(pr, AB_PR_LINV, AbView{c0d, B, 0, nlimbs, AB_LAYOUT_PLANAR}, st);
            ab_launch_tunnel_gather(ps, p0, c0d, per_r, nr, t->d_dec, t->d_dec_sgn, B, K, nlimbs, st);
            ab_launch_prime_op(ps, AB_PR_L, AbView{p0, B * K, 0, nlimbs, AB_LAYOUT_PLANAR}, st);
            ab_crt_dispatch(ps, AbView{p0, B * K, 0, nlimbs, AB_LAYOUT_PLANAR}, false, st);
            // evalLin: c0' = sum_k p0_k * flin_k
            ab_launch_lin_ip(ps, lin, per_s, p0, K * per_s, flin, B, nlimbs, (int32_t)K, st);
            // c1: split over E' in the powerful basis (coeffsPow), embed (embedPow), decompose, reduce, CRT
            if (fused_ok(ps, nlimbs, dig, cin)) {
                // one launch: gather (coeffs . embed), gadget digit, reduction and forward CRT of O_s'
                AbMxFused fx{};
                fx.load_mode = AB_MX_LOAD_DIGITS;
                fill_digits(ps, nlimbs, base, fx, false);
                fx.src = cin + per_r; fx.src_ustride = 2 * per_r; fx.n_src = (int32_t)nr; fx.gather = t->d_pow; fx.gK = (int32_t)K;
                ab_launch_crt_mixed_fused(ps, AbView{dig, B * K * nd, 0, nlimbs, AB_LAYOUT_PLANAR}, false, &fx, st);
            } else {
                ab_launch_tunnel_gather(ps, p1, cin + per_r, 2 * per_r, nr, t->d_pow, nullptr, B, K, nlimbs, st);
                ab_launch_decompose(ps, dig, p1, B * K, nlimbs, nlimbs, base, st);
                ab_crt_dispatch(ps, AbView{dig, B * K * nd, 0, nlimbs, AB_LAYOUT_PLANAR}, false, st);
            }
            // (c0'', c1'') = sum_{k,d} dig_{k,d} * hint_{k,d};  then c0'' += c0'
            ab_launch_hint_ip_m(ps, o, 2 * per_s, dig, K * nd * per_s, hm, nullptr, 0, B, nlimbs, (int32_t)(K * nd), st);
            ab_launch_add_rows(ps, o, 2 * per_s, lin, per_s, B, nlimbs, st);
        }
    });
}

// ------------------------------------------------------------------------------------ PT2CT mul_ as one entry
// modSwitch (to l_out) . keySwitchQuad hint . modSwitch (to l_hint) . (*)   -- Crypto/Alchemy/Interpreter/PT2CT.hs:172-177.
// Lol's (*) returns an LSD ciphertext and modSwitch / keySwitchQuadCirc first bring it to MSD form (toMSD: every polynomial times
// p^-1 mod q_l; ModSwitchCtx_ carries ToSDCtx, Eval.hs:124): the caller passes those per-limb scalars as `prescale` (null = the
// inputs are taken as they are, e.g. already pre-scaled) and keeps the (k, l) bookkeeping on its side.
// Small power-of-two rings with TrivGad run as ONE kernel with every intermediate in shared memory
// (ab_mulks_pow2.cu); everything else as the sequence of the entries above, intermediates in `work`.
static int64_t mks_inter_words(const ab_plan *p, int64_t batch, int32_t lin, int32_t lh) {
    return batch * p->n * (3 * (int64_t)lin + 3 * (int64_t)lh + 2 * (int64_t)lh);
}
extern "C" int64_t ab_mul_key_switch_work_bytes(const ab_plan *p, int64_t batch, int32_t lin, int32_t lh, int32_t lout, int32_t base) {
    if (!p || batch < 0 || lin < 1 || lh < 1 || lout < 1 || lin > p->L || lh > p->L || lout > p->L) return -1;
    if (ab_mulks_n256_supported(p, lin, lh, lout, base) || ab_mulks_pow2_supported(p, lin, lh, lout, base)) return 16;
    const int64_t w1 = ab_mod_switch_work_bytes(p, batch, 3, lin, lh), w2 = ab_key_switch_work_bytes(p, batch, lh, base),
                  w3 = ab_mod_switch_work_bytes(p, batch, 2, lh, lout);
    if (w1 < 0 || w2 < 0 || w3 < 0) return -1;
    return mks_inter_words(p, batch, lin, lh) * 8 + std::max(w1, std::max(w2, w3)) + 64;
}
extern "C" int ab_mul_key_switch_dev(const ab_plan *p, int64_t *out2, const int64_t *a2, const int64_t *b2, const int64_t *hint, int64_t batch,
                                     int32_t lin, int32_t lh, int32_t lout, int32_t base, const int64_t *prescale, uint32_t dec_mask, void *work, void *stream) {
    return ab_guard([&] {
        AB_REQUIRE(p && ((out2 && a2 && b2 && hint) || batch == 0), AB_ERR_ARG, "null argument");
        AB_REQUIRE(p->has_crt, AB_ERR_NO_ROOT, "plan has no CRT tables");
        AB_REQUIRE(batch >= 0 && lin >= 1 && lh >= 1 && lout >= 1 && lin <= p->L && lh <= p->L && lout <= p->L, AB_ERR_ARG, "bad batch or limb counts");
        if (batch == 0) return;
        if (ab_mulks_n256_supported(p, lin, lh, lout, base)) {
            ab_launch_mulks_n256(p, out2, a2, b2, hint, batch, (cudaStream_t)stream, prescale);      // examples/Arithmetic.hs: one warp per op
            return;
        }
        if (ab_mulks_pow2_supported(p, lin, lh, lout, base)) {
            ab_launch_mulks_pow2(p, out2, a2, b2, hint, batch, lin, lh, lout, (cudaStream_t)stream, prescale);      // no odd prime: Dec = Pow
            return;
        }
        AB_REQUIRE(work != nullptr, AB_ERR_ARG, "null scratch buffer");
        const int64_t n = p->n;
        int64_t *c3 = (int64_t *)work, *c3h = c3 + batch * 3 * lin * n, *c2h = c3h + batch * 3 * lh * n;
        void *w = c2h + batch * 2 * lh * n;
        auto must = [](int rc) { if (rc != AB_OK) throw ab_error{rc, ab_last_error()}; };
        // (*) yields an LSD ciphertext; modSwitch starts with toMSD = every polynomial times p^-1 mod q_l (`prescale`), folded into the product
        if (lh >= lin) {
            ab_launch_ct_mul_up(p, c3h, a2, b2, batch, lin, lh, (cudaStream_t)stream, prescale);      // product, toMSD and scale-up in one pass
        } else {
            ab_launch_ct_mul_up(p, c3, a2, b2, batch, lin, lin, (cudaStream_t)stream, prescale);
            must(ab_mod_switch_dev(p, c3h, c3, batch, 3, lin, lh, dec_mask, w, stream));
        }
        // after a scale-up the appended limbs of the product are exactly zero: their gadget digits vanish and are skipped
        static const bool no_skip = getenv("AB_NO_ZERO_SKIP") && atoi(getenv("AB_NO_ZERO_SKIP")) != 0;      // A/B measurements only
        must(key_switch_quad(p, c2h, c3h, hint, batch, lh, base, w, stream, (lh > lin && !no_skip) ? lin : lh));
        must(ab_mod_switch_dev(p, out2, c2h, batch, 2, lh, lout, dec_mask, w, stream));
    });
}
