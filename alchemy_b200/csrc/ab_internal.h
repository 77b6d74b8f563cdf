// ab_internal.h -- shared host-side declarations of the B200 tensor backend (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/alchemy_b200.h"

#define AB_MAX_LIMBS 8
#define AB_MAX_AXES 12

// ---------------------------------------------------------------------------------------------
// error plumbing: every ABI entry is `return guard([&]{ ... });`
void ab_set_error(const std::string &msg);
struct ab_error {
    int code;
    std::string msg;
};
#define AB_CHECK_CUDA(expr)                                                                       \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess)                                                                    \
            throw ab_error{AB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)};      \
    } while (0)
#define AB_REQUIRE(cond, code, msg)                                                               \
    do {                                                                                          \
        if (!(cond)) throw ab_error{(code), (msg)};                                               \
    } while (0)

template <class F>
static inline int ab_guard(F &&f) {
    try {
        f();
        return AB_OK;
    } catch (const ab_error &e) {
        ab_set_error(e.msg);
        return e.code;
    } catch (const std::exception &e) {
        ab_set_error(e.what());
        return AB_ERR_ARG;
    } catch (...) {
        ab_set_error("unknown failure");
        return AB_ERR_ARG;
    }
}

void ab_count_launch(uint64_t n = 1);
void ab_check_launch(const char *what);   // cudaGetLastError after a launch; throws ab_error

// ---------------------------------------------------------------------------------------------
// host number theory (ab_plan.cpp)
namespace abh {
uint64_t mulmod(uint64_t a, uint64_t b, uint64_t q);
uint64_t powmod(uint64_t b, uint64_t e, uint64_t q);
uint64_t invmod(uint64_t a, uint64_t q);   // q prime
bool is_prime(uint64_t q);
uint64_t smallest_generator(uint64_t q);
int64_t ipow(int64_t b, int e);
int64_t digit_rev(int p, int e, int64_t j);
static inline uint32_t shoup(uint64_t w, uint64_t q) { return (uint32_t)((w << 32) / q); }
}   // namespace abh

// ---------------------------------------------------------------------------------------------
// device-side parameter blocks (plain structs passed by value to kernels)
struct uint2_ {
    uint32_t x, y;
};   // (w, floor(w*2^32/q)) Shoup pair; same layout as CUDA uint2

// ---------------------------------------------------------------------------------------------
// power-of-two kernel geometry, shared by the table builder (ab_plan.cpp) and the kernels (ab_ntt_pow2.cu).
// A transform of n = 2^logn points is cut into windows of `loge` butterfly stages held in registers;
// window w starts at stage S(w) and skips its first LB0(w) stages (only the last window can overlap).
// Twiddles are stored window-packed: for window w, entry k = 2^lb + ip - koff(w) of column `low` sits at
//   base(w) + ((k/2) * 2^S + low) * 2 + (k & 1)        (uint2 units)
// so that one 16-byte load per lane fetches two consecutive entries and lanes are contiguous.
#if defined(__CUDACC__)
#define AB_HD __host__ __device__
#else
#define AB_HD
#endif
AB_HD constexpr int ab_pow2_loge(int logn) { return logn == 5 ? 5 : logn == 6 ? 3 : logn <= 8 ? 4 : 5; }
AB_HD constexpr int ab_pow2_nw(int logn, int loge) { return (logn + loge - 1) / loge; }
AB_HD constexpr int ab_pow2_s(int logn, int loge, int w) { return (w * loge + loge <= logn) ? w * loge : logn - loge; }
AB_HD constexpr int ab_pow2_lb0(int logn, int loge, int w) { return w * loge - ab_pow2_s(logn, loge, w); }
AB_HD constexpr int ab_pow2_koff(int logn, int loge, int w) { return ab_pow2_lb0(logn, loge, w) == 0 ? 0 : (1 << ab_pow2_lb0(logn, loge, w)); }
AB_HD constexpr int64_t ab_pow2_wsize(int logn, int loge, int w) {   // uint2 entries of window w
    return (int64_t)((1 << loge) - ab_pow2_koff(logn, loge, w)) << ab_pow2_s(logn, loge, w);
}
AB_HD constexpr int64_t ab_pow2_base(int logn, int loge, int w) {
    int64_t b = 0;
    for (int i = 0; i < w; ++i) b += ab_pow2_wsize(logn, loge, i);
    return b;
}

struct AbLimbDev {
    uint32_t q;
    uint32_t ninv, ninv_s;       // pow2 path: n^-1 (Shoup pair) folded into the last inverse stage
    uint64_t mu;                 // floor(2^64 / q) for Barrett reduction of 62-bit products
    const uint2_ *tw_fwd;        // pow2 path: merged negacyclic twiddles, window-packed (see above)
    const uint2_ *tw_inv;        // pow2 path: their inverses
    uint2_ tw0_fwd[32], tw0_inv[32];   // host copies of window 0 (same for every thread): passed as kernel constants
    const uint2_ *tw_fwd_flat = nullptr; // pow2 path: the same twiddles unpacked, entry [2^s + l] (ab_mulks_pow2.cu)
    const uint2_ *tw_inv_flat = nullptr;
    const uint2_ *gen_fwd;       // generic path: per-limb table block (matrices, twiddles)
    const uint2_ *gen_inv;
    const uint2_ *g_crt;         // g at every CRT row (Shoup pairs), n entries
    const uint2_ *ginv_crt;
    uint32_t pinv[AB_MAX_AXES];  // p^-1 mod q per axis (divG)
    uint32_t pinv_s[AB_MAX_AXES];
};

// one step of the generic mixed-radix transform (device array, shared by all limbs).
// Work item v of a step: tb = v / estride, off = v % estride, base = tb*P*estride + off; its P
// elements sit at base + k*estride.
struct AbGenStep {
    int32_t kind;       // 0 = dense PxP matvec, scale row = tb % h (CRT_p steps)
                        // 2 = dense PxP matvec, scale row = off / sdiv (DFT_p stages)
                        // 1 = radix-2 butterfly stage, twiddle row = off / sdiv
    int32_t P;          // vector length (p-1 for CRT_p, p for DFT_p, 2 for radix-2)
    int32_t estride;    // distance between the P elements of a work item
    int32_t h;          // kind 0: p^(e-1); kinds 1,2: sub-transform size (rows of the twiddle table)
    int32_t sdiv;       // kinds 1,2: twiddle row = off / sdiv
    int32_t mode;       // 0 no scaling, 1 scale inputs, 2 scale outputs (table entry [row*P + k])
    int32_t mat_off;    // offset (Shoup pairs) of the PxP matrix in the per-limb table block
    int32_t sc_off;     // offset of the scale / twiddle table
    int32_t nwork;      // work items in this step (n / P)
    int32_t flags;      // kind 1 inverse: bit0 = multiply output 0 by table entry [h] (axis 2^-(e-1))
};

// ---------------------------------------------------------------------------------------------
// fast mixed-radix program (ab_crt_mixed.cu): every step is one pass over the n residues of a transform held
// in shared memory.  Work item v of a step: tb = v / estride, off = v % estride, base = tb*span*estride + off;
// its `span` elements sit at base + k*estride.
struct AbMxStep {
    int32_t kind;        // 0 = dense span x span mat-vec (Montgomery constants), 1 = group of G radix-2 stages (Shoup twiddles),
                         // 2 = CRT_p of a prime axis (span = p - 1) through its conjugate symmetry: half the multiplications (ab_mixed_plan.cpp)
    int32_t span;        // mat-vec: P;  radix-2: 2^G
    int32_t g;           // radix-2: G (1..3)
    int32_t estride;
    int32_t nwork;       // n / span
    uint32_t div_magic;  // v / estride == umulhi(v, div_magic)           (0: estride == 1)
    int32_t row_mode;    // matrix / twiddle row: 0 = none, 1 = tb % rows, 2 = off / sdiv
    int32_t rows;
    uint32_t row_magic;  // mode 1: tb / rows == umulhi(tb, row_magic); mode 2: off / sdiv == umulhi(off, row_magic)
    int32_t sdiv;
    int32_t tab_off;     // uint32 words into the limb's table block, narrow layout (q < 2^30)
    int32_t tab_off_w;   // same for the wide layout (2^30 <= q < 2^31: the constants are centred representatives stored as int32; same sizes)
    int32_t flags;       // radix-2 inverse: bit0 = this group ends the axis: fold 2^-(e-1) into the last stage
    int32_t vec;         // 1 or 2 adjacent columns per work item (2: estride even, 8-byte shared-memory accesses);
                         // div_magic then divides by estride / vec
};
#define AB_MX_MAX_STEPS 24

// fused load / store of the mixed-radix CRT kernel (key switch, tunnel, modSwitch; ab_she.cpp)
enum { AB_MX_LOAD_PLAIN = 0, AB_MX_LOAD_DIGITS = 1, AB_MX_LOAD_SRC = 2, AB_MX_LOAD_GARNER = 3, AB_MX_LOAD_GATHER = 4 };
#define AB_MX_MAX_DROP 4
enum { AB_MX_STORE_PLAIN = 0, AB_MX_STORE_SUBSCALE = 1 };
#define AB_MX_MAX_DIGITS 64
struct AbMxFused {
    int32_t load_mode, store_mode;
    // AB_MX_LOAD_DIGITS: transform (unit * nd + d, limb l') = CRT of digit d of src[unit / gK] reduced mod q_l'
    // AB_MX_LOAD_SRC:    transform (unit, limb) read from src + unit * src_ustride + (src_limb0 + limb) * n
    // AB_MX_LOAD_GATHER: transform (unit, limb)[j] = src[unit / gK][limb][gather[unit % gK][j]]  (coeffs of the tunnel's c0 part; -1 = zero)
    int32_t nd, gbase;               // gbase 0 = TrivGad
    const int64_t *src;              // powerful basis, planar [unit][limb][n_src]
    int64_t src_ustride;
    int32_t n_src, src_limb0;
    const int32_t *gather;           // optional [gK][n]: index into the n_src source coefficients (-1 = zero)
    int32_t gK, pad0;
    const int64_t *triv_src;         // TrivGad: CRT-basis copy of the source (digit of limb l mod q_l needs no transform)
    int64_t triv_ustride;
    // digit i of limb l (k digits, base b, hb = floor(b/2)): with H = hb (b^(k-1) - 1)/(b - 1), the balanced digits of
    // Lol's iterated divModCent are the ordinary base-b digits of c + H shifted by hb:
    //   d_i = (floor((c + H) / b^i) mod b) - hb  (i < k-1),   d_(k-1) = floor((c + H) / b^(k-1))
    struct { uint32_t q_src, pw, H; int16_t limb, last; } dig[AB_MX_MAX_DIGITS];
    // AB_MX_LOAD_GARNER: several RNS limbs dropped in ONE rescale (modSwitch from L to L - gar_k limbs).  src holds the
    // dropped limbs in the powerful basis, [unit][gar_k][n] in DROP order (i = 0 is the last limb).  Per coefficient the
    // centred digits of the sequential drops are r_0 = centre(v_0), v_t <- (v_t - r_0) q_0^-1 mod q_t, r_1 = centre(v_1), ...
    // and the value to subtract in every kept limb is x = r_0 + q_0 (r_1 + q_1 (...)) mod q  (Horner): dropping limbs one at a
    // time gives the same integers, because q_0 r_1 + r_0 is again a centred remainder (|.| <= (q_0 q_1 - 1) / 2).
    int32_t gar_k, pad1;
    uint32_t gar_q[AB_MX_MAX_DROP];
    uint32_t gar_inv[AB_MX_MAX_DROP][AB_MX_MAX_DROP], gar_inv_s[AB_MX_MAX_DROP][AB_MX_MAX_DROP];   // [i][t] = q_i^-1 mod q_t, t > i
    // Dec <-> Pow folded into the transform: pre_l applies L (decoding -> powerful) to the loaded residues before a forward
    // program, post_linv applies L^-1 after an inverse program (Lol `l` / `lInv`: running sums / differences along the
    // p-1 axis of every odd prime); the axes are filled in by the launcher
    int32_t pre_l, post_linv, n_odd, pad2;
    // l_period > 1: the L passes run only for units u with bit ((u + l_phase) % l_period) of l_mask set (modSwitch of a
    // ciphertext whose polynomials are rescaled in different bases: Lol rescaleLinearMSD takes c0 in Dec, c1 in Pow [RECALL])
    int32_t l_period, l_mask, l_phase, pad3;
    int64_t dst_ustride;             // 0: units are contiguous in the destination; else int64 words between consecutive units
    struct { int32_t d, rts; } odd[AB_MAX_AXES];
    // AB_MX_STORE_SUBSCALE: y = (sub_in - x) * scale[limb]
    const int64_t *sub_in;
    int64_t sub_ustride;
    uint32_t scale[AB_MAX_LIMBS], scale_s[AB_MAX_LIMBS];
};

// fused gadget inner product (ab_crt_mixed.cu, k_crt_mixed_mac): one CTA per (unit, limb) walks over the K * nd
// gadget digits of the unit -- digit extraction, forward CRT and the multiply-accumulate against the prepared
// hint all happen in shared memory; only the two result polynomials leave the SM.
struct AbMxMac {
    const void *hm;              // prepared hint [K * nd][L][n] of uint2 = (h_0 2^32 mod q, h_1 2^32 mod q)  (ab_launch_hint_prep2)
    int32_t K, parts;            // tunnel: dimension of O_r' over O_e' (digit loop runs over K * nd); key switch: 1.
                                 // parts > 1: the digit list of a unit is cut into `parts` CTAs (more, shorter CTAs: fuller
                                 // last wave); partial sums stay in `acc` [unit][part][L][n] and k_mac_combine adds them up
    const int64_t *add;          // optional [unit][2][L][n] added to both components (key switch: c0, c1)
    int64_t add_ustride;
    const int64_t *add0;         // optional [unit][L][n] added to component 0 (tunnel: the evalLin part)
    int64_t add0_ustride;
    const uint32_t *bits_lut[AB_MAX_LIMBS];   // per limb: subset-sum tables of the first step for binary digits (null: generic path)
    int32_t bits_p, bits_h;
    // replicated mode (tunnel): the digits live in the SUB-ring O_e' (the plan handed to the kernel, n = phi(e')) and the
    // inner product runs over the n_out = phi(s') CRT rows of O_s': CRT_s'(embed x) = embedCRT(CRT_e'(x)) is a replication of
    // the small transform through rep_idx[j] = CRT row of O_e' below row j of O_s' (Lol `embedCRT`), so every digit
    // transform is phi(s') / phi(e') times smaller.  n_out = 0: plain mode (hint, sums and output over the plan's own n).
    const int32_t *rep_idx;
    int32_t n_out, group;        // group: digit transforms held in shared memory per inner-product pass (1..8; filled in by the launcher)
    int32_t rep4, pad_rep;       // replicated mode: every aligned group of four output rows reads four CONSECUTIVE sub-ring rows starting at a
                                 // multiple of four (true whenever the innermost axis is shared and a multiple of 4 long): one 16-byte shared load
                                 // per group instead of four 4-byte loads, which sat four lanes on every bank (3.3 x the ideal wavefronts measured)
    void *acc;                   // scratch of ab_mixed_mac_scratch_words() int64 words (may be null when that is 0)
    int64_t *tbuf;               // optional scratch of [unit][parts][L][n] int64 words (used as uint32: the source residues gathered once per (k, source limb))
    int64_t *out;                // [unit][2][L][n]
    int64_t out_ustride;
};

struct AbAxis {
    int32_t p, e;
    int64_t pp, phi, mp;    // p^e, phi(p^e), p^(e-1)
    int64_t lts, rts;       // Kronecker sizes to the left / right
};

struct ab_plan {
    int64_t m = 0, n = 0;
    int32_t L = 0;
    bool has_crt = false, pow2 = false;
    int32_t logn = 0;
    std::vector<AbAxis> axes;
    std::vector<int64_t> q, root;
    std::vector<AbLimbDev> limb;          // host copy of device parameter blocks
    // generic transform program
    std::vector<AbGenStep> steps_fwd, steps_inv;
    AbGenStep *d_steps_fwd = nullptr, *d_steps_inv = nullptr;
    std::vector<void *> owned;            // device allocations to free
    int device = 0;
    int sm_count = 148;
    // fast mixed-radix program (moduli below 2^30 only; see ab_mixed_plan.cpp)
    std::vector<AbMxStep> mx_fwd, mx_inv;
    std::vector<const uint32_t *> mx_tab_fwd, mx_tab_inv;   // per limb, device
    int32_t mx_words_fwd[2] = {0, 0}, mx_words_inv[2] = {0, 0};   // table block size (uint32 words): narrow, wide layout
    // binary-digit first step (gadget base 2): when the first forward step is the twiddle-free (p-1)-point CRT_p of the last
    // axis, the step applied to digits in {-1, 0} is a subset sum of matrix columns: per limb two lookup tables
    // lut[half][row][subset] = -(sum of M[row][c] over the columns c of that half whose bit is set) mod q, half = low / high
    // ceil(P/2) columns.  mx_bits_p = P (0: not available), mx_bits_h = ceil(P/2); device pointer per limb.
    int32_t mx_bits_p = 0, mx_bits_h = 0;
    std::vector<const uint32_t *> mx_bits_lut;
    std::vector<uint8_t> mx_ok;                             // per limb: 0 = not supported, 1 = narrow (q < 2^30), 2 = wide
    // host copies kept for tests / ext construction
    std::vector<int64_t> mhat_inv;        // per limb
};

struct ab_ext {
    const ab_plan *lo = nullptr, *hi = nullptr;
    int64_t n_lo = 0, n_hi = 0, rel = 0;
    int32_t *d_ext_pow = nullptr;      // [n_lo]   index into hi of pow/dec basis element j of lo
    int32_t *d_pow_to_lo = nullptr;    // [n_hi]   inverse map (or -1)
    int32_t *d_base_crt = nullptr;     // [n_hi]   CRT row of lo below CRT row of hi
    int32_t *d_above_crt = nullptr;    // [n_lo][rel] CRT rows of hi above each row of lo
    int32_t *d_coeff = nullptr;        // [rel][n_lo] coeffs index table
    // embedDec is a signed gather: out_hi[j'] = sgn[j'] * in_lo[idx[j']]  (idx < 0: zero)
    int32_t *d_edec_idx = nullptr;
    int8_t *d_edec_sgn = nullptr;
    // twaceCRT weights per limb: w[i'] = (mhat/mhat') * g'(i') / g(base(i'))  (Shoup pairs), [L][n_hi]
    uint2_ *d_tw_weight = nullptr;
    std::vector<void *> owned;
};

// ---------------------------------------------------------------------------------------------
// kernel launchers (one .cu each)
struct AbView {             // a batch of ring elements on the device
    int64_t *y;
    int64_t batch;
    int32_t limb0, nlimbs;
    int layout;
};

void ab_launch_crt_pow2(const ab_plan *p, const AbView &v, bool inverse, cudaStream_t st);
bool ab_pow2_supported(const ab_plan *p);
void ab_launch_crt_generic(const ab_plan *p, const AbView &v, bool inverse, cudaStream_t st);
void ab_build_mixed(ab_plan *pl);                                 // ab_mixed_plan.cpp
bool ab_mixed_supported(const ab_plan *p, const AbView &v);       // ab_crt_mixed.cu
void ab_launch_crt_mixed(const ab_plan *p, const AbView &v, bool inverse, cudaStream_t st);
void ab_launch_crt_mixed_fused(const ab_plan *p, const AbView &v, bool inverse, const AbMxFused *fx, cudaStream_t st);
// units x nlimbs CTAs; fx describes the digit source (load_mode AB_MX_LOAD_DIGITS); plan p = the ring the digits live in
// whole PT2CT mul_ pipeline in one kernel for small power-of-two rings with TrivGad (ab_mulks_pow2.cu)
bool ab_mulks_pow2_supported(const ab_plan *p, int32_t lin, int32_t lh, int32_t lout, int32_t base);
void ab_launch_mulks_pow2(const ab_plan *p, int64_t *out2, const int64_t *a2, const int64_t *b2, const int64_t *hint, int64_t batch,
                          int32_t lin, int32_t lh, int32_t lout, cudaStream_t st, const int64_t *prescale = nullptr);
// the same pipeline for the limb schedule of examples/Arithmetic.hs on n = 256 (2 -> 2 -> 1 limbs, TrivGad): one warp per op, transforms in registers (ab_mulks_n256.cu)
bool ab_mulks_n256_supported(const ab_plan *p, int32_t lin, int32_t lh, int32_t lout, int32_t base);
void ab_launch_mulks_n256(const ab_plan *p, int64_t *out2, const int64_t *a2, const int64_t *b2, const int64_t *hint, int64_t batch,
                          cudaStream_t st, const int64_t *prescale = nullptr);
bool ab_mixed_mac_supported(const ab_plan *p, int32_t nlimbs);
int64_t ab_mixed_mac_scratch_words(const ab_plan *p, int64_t units, int32_t nlimbs);
int32_t ab_mixed_mac_parts(const ab_plan *p, int64_t units, int32_t nlimbs, int32_t total_digits);   // how many CTAs per unit pay off
void ab_launch_crt_mixed_mac(const ab_plan *p, int64_t units, int32_t nlimbs, const AbMxFused &fx, const AbMxMac &mac, cudaStream_t st);

enum { AB_EW_MUL = 0, AB_EW_ADD = 1, AB_EW_SUB = 2, AB_EW_NEG = 3, AB_EW_SCALAR = 4, AB_EW_GCRT = 5, AB_EW_GINVCRT = 6 };
void ab_launch_elementwise(const ab_plan *p, int op, const AbView &v, const int64_t *x, int64_t x_batch,
                           const int64_t *scalars_host, cudaStream_t st);
enum { AB_PR_L = 0, AB_PR_LINV = 1, AB_PR_GPOW = 2, AB_PR_GINVPOW = 3, AB_PR_GDEC = 4, AB_PR_GINVDEC = 5 };
void ab_launch_prime_op(const ab_plan *p, int op, const AbView &v, cudaStream_t st);

void ab_launch_gather(int64_t *out, const int64_t *in, const int32_t *idx, int64_t rows, int64_t n_out, int64_t n_in,
                      cudaStream_t st);   // out[r][j] = idx[j] >= 0 ? in[r][idx[j]] : 0
void ab_launch_layout(int64_t *out, const int64_t *in, int64_t batch, int32_t L, int64_t n, bool to_planar, cudaStream_t st);
