/*
 * alchemy_b200.h -- C ABI of the B200-native tensor backend for ALCHEMY / Lol.
 *
 * This library replaces the `lol-cpp` instance of Lol's `class Tensor` -- the backend the
 * reference selects with the type tag `CT` at examples/Arithmetic.hs:19,23,
 * examples/HomomRLWR.hs:24,47 and examples/Tunnel.hs:20,41 (cabal dependency `lol-cpp`,
 * alchemy.cabal:74,93,117).  Nothing under Crypto/Alchemy names a backend; every hot-path
 * operation is reached from the right-hand sides of the `E` interpreter
 * (Crypto/Alchemy/Interpreter/Eval.hs:58-67,95,129-134) through Lol's Cyc/SymmSHE layers.
 *
 * The upstream C entry points (`tensorCRTRq`, `tensorCRTInvRq`, `tensorLRq`, `tensorGPowRq`,
 * `mulRq`, ... in lol-cpp) are NOT in /root/reference; their names and signatures below are
 * the surveyor's recollection and are marked [RECALL].
 *
 * Rules of the ABI
 *   - plain C: pointers, sizes, int status codes.  0 = AB_OK; no exception or abort crosses it.
 *   - re-entrant: no global mutable state except the per-thread last-error string.
 *   - residues are int64_t in [0,q) (Lol `ZqBasic q Int64`, examples/Common.hs:35); q prime, q < 2^31.
 *   - `*_dev` entries take DEVICE pointers and a CUDA stream (as void*, 0 = default stream) and
 *     are asynchronous; `*_host` entries take HOST pointers, copy in, run, copy out, and return
 *     after the result is in the host buffer (the behaviour of a `foreign import ccall` today).
 *     The `*_dev` compute entries neither synchronise nor allocate (scratch comes from the caller, `*_work_bytes`), so a chain of
 *     them can be captured into a CUDA graph and replayed (call each once before capturing: kernel attributes are set lazily).
 *     Exceptions: ab_wire_pack_dev / ab_wire_unpack_dev / ab_wire_read_header return after a device-to-host read.
 *   - buffers are owned by the caller; plans own their device tables (create/destroy).
 *   - there is no CPU fallback: without a CUDA device every compute entry returns AB_ERR_CUDA.
 *
 * Layouts of a batch of ring elements with L limbs (RNS tuple size) and n = phi(m) coefficients:
 *   AB_LAYOUT_PLANAR      y[(b*L + l)*n + j]     the backend's device-resident format
 *   AB_LAYOUT_INTERLEAVED y[(b*n + j)*L + l]     Lol's Storable tuple vectors ([RECALL]: lol-cpp `tupSize`)
 */
#ifndef ALCHEMY_B200_H
#define ALCHEMY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    AB_OK = 0,
    AB_ERR_ARG = 1,       /* bad argument (null, size, layout, modulus not supported) */
    AB_ERR_NO_ROOT = 2,   /* m does not divide q-1: no CRT over Z_q (Lol: crtFuncs = Nothing) */
    AB_ERR_CUDA = 3,      /* CUDA runtime error or no device; see ab_last_error() */
    AB_ERR_UNSUPPORTED = 4
};

enum { AB_LAYOUT_PLANAR = 0, AB_LAYOUT_INTERLEAVED = 1 };

/* lol-cpp `PrimeExponent {hShort_t prime; hShort_t exponent;}` [RECALL]; widened to 32 bits */
typedef struct { int32_t prime; int32_t exponent; } ab_prime_exp;

typedef struct ab_plan ab_plan;         /* one cyclotomic index m with L moduli: all device tables */
typedef struct ab_ext ab_ext;           /* a pair m | m' of plans: twace/embed/coeffs index tables */

const char *ab_last_error(void);        /* message of the last failure on this thread */
int ab_device_count(void);              /* 0 when no CUDA device is usable */
const char *ab_version(void);

/* ---- plans -------------------------------------------------------------------------------
 * `m` = prod pe[i].prime^pe[i].exponent (any order; sorted internally, primes ascending).
 * `roots`: optional omega_m per limb (what lol-cpp receives as `ru` from Haskell); NULL = use
 * omega_m = g^((q-1)/m) with g the smallest generator of Z_q^* (Lol `principalRootUnity`).
 * Fails with AB_ERR_NO_ROOT if some q has no m-th root of unity: then only the non-CRT entries
 * (L, G in Pow/Dec, twace/embed Pow/Dec, pointwise) may be used on a plan created with
 * ab_plan_create_nocrt.  A nocrt plan also accepts composite moduli 2 <= q < 2^31 (plaintext rings Z_p, p = 2^k
 * in examples/HomomRLWR.hs): divG then needs every odd prime of m invertible mod q (AB_ERR_UNSUPPORTED otherwise). */
int ab_plan_create(ab_plan **out, const ab_prime_exp *pe, int32_t num_pe,
                   const int64_t *moduli, int32_t num_limbs, const int64_t *roots);
int ab_plan_create_nocrt(ab_plan **out, const ab_prime_exp *pe, int32_t num_pe,
                         const int64_t *moduli, int32_t num_limbs);
void ab_plan_destroy(ab_plan *plan);
int64_t ab_plan_m(const ab_plan *plan);
int64_t ab_plan_n(const ab_plan *plan);          /* phi(m) */
int32_t ab_plan_limbs(const ab_plan *plan);
int64_t ab_plan_root(const ab_plan *plan, int32_t limb);   /* omega_m in use for that limb */

/* ---- Tensor methods, device buffers, in place on y, batched ----------------------------------
 * Each replaces the lol-cpp symbol named at its right ([RECALL]); `batch` ring elements at once.
 * `limb0`/`nlimbs` select a contiguous sub-range of the plan's moduli for buffers that hold only
 * some limbs (limb l of the buffer uses plan modulus limb0 + l; buffer tuple size = nlimbs). */
int ab_crt_dev(const ab_plan *p, int64_t *y, int64_t batch, int32_t limb0, int32_t nlimbs, int layout, void *stream);       /* tensorCRTRq */
int ab_crt_inv_dev(const ab_plan *p, int64_t *y, int64_t batch, int32_t limb0, int32_t nlimbs, int layout, void *stream);   /* tensorCRTInvRq (includes the mhat^-1 scaling) */
int ab_l_dev(const ab_plan *p, int64_t *y, int64_t batch, int32_t limb0, int32_t nlimbs, int layout, void *stream);         /* tensorLRq     Dec -> Pow */
int ab_l_inv_dev(const ab_plan *p, int64_t *y, int64_t batch, int32_t limb0, int32_t nlimbs, int layout, void *stream);     /* tensorLInvRq  Pow -> Dec */
int ab_mul_g_pow_dev(const ab_plan *p, int64_t *y, int64_t batch, int32_t limb0, int32_t nlimbs, int layout, void *stream); /* tensorGPowRq */
int ab_mul_g_dec_dev(const ab_plan *p, int64_t *y, int64_t batch, int32_t limb0, int32_t nlimbs, int layout, void *stream); /* tensorGDecRq */
int ab_div_g_pow_dev(const ab_plan *p, int64_t *y, int64_t batch, int32_t limb0, int32_t nlimbs, int layout, void *stream); /* tensorGInvPowRq (field division: always succeeds mod q) */
int ab_div_g_dec_dev(const ab_plan *p, int64_t *y, int64_t batch, int32_t limb0, int32_t nlimbs, int layout, void *stream); /* tensorGInvDecRq */
int ab_mul_g_crt_dev(const ab_plan *p, int64_t *y, int64_t batch, int32_t limb0, int32_t nlimbs, int layout, void *stream); /* tensorGCRTRq */
int ab_div_g_crt_dev(const ab_plan *p, int64_t *y, int64_t batch, int32_t limb0, int32_t nlimbs, int layout, void *stream); /* tensorGInvCRTRq */

/* zipWithT (*), (+), (-) and scalar multiplication on Z_q tuples; x may be broadcast over the
 * batch with x_batch == 1 (a public polynomial or a hint against a batch of ciphertexts). */
int ab_mul_dev(const ab_plan *p, int64_t *y, const int64_t *x, int64_t batch, int64_t x_batch, int32_t limb0, int32_t nlimbs, int layout, void *stream); /* mulRq */
int ab_add_dev(const ab_plan *p, int64_t *y, const int64_t *x, int64_t batch, int64_t x_batch, int32_t limb0, int32_t nlimbs, int layout, void *stream);
int ab_sub_dev(const ab_plan *p, int64_t *y, const int64_t *x, int64_t batch, int64_t x_batch, int32_t limb0, int32_t nlimbs, int layout, void *stream);
int ab_neg_dev(const ab_plan *p, int64_t *y, int64_t batch, int32_t limb0, int32_t nlimbs, int layout, void *stream);
/* y[., l, .] *= scalars[l]   (scalars: HOST array of nlimbs residues) */
int ab_scalar_mul_dev(const ab_plan *p, int64_t *y, const int64_t *scalars, int64_t batch, int32_t limb0, int32_t nlimbs, int layout, void *stream);

/* ---- extension rings  m | m'  (Lol twacePowDec / embedPow / embedDec / twaceCRT / embedCRT / coeffs) --
 * `lo` is the plan of O_m, `hi` the plan of O_m'; both must carry the same moduli.  Out-of-place;
 * PLANAR layout only; `nlimbs` leading limbs of both plans. */
int ab_ext_create(ab_ext **out, const ab_plan *lo, const ab_plan *hi);
void ab_ext_destroy(ab_ext *ext);
int ab_twace_pow_dec_dev(const ab_ext *e, int64_t *out_lo, const int64_t *in_hi, int64_t batch, int32_t nlimbs, void *stream);
int ab_embed_pow_dev(const ab_ext *e, int64_t *out_hi, const int64_t *in_lo, int64_t batch, int32_t nlimbs, void *stream);
int ab_embed_dec_dev(const ab_ext *e, int64_t *out_hi, const int64_t *in_lo, int64_t batch, int32_t nlimbs, void *stream);
int ab_twace_crt_dev(const ab_ext *e, int64_t *out_lo, const int64_t *in_hi, int64_t batch, int32_t nlimbs, void *stream);
int ab_embed_crt_dev(const ab_ext *e, int64_t *out_hi, const int64_t *in_lo, int64_t batch, int32_t nlimbs, void *stream);
/* coeffs: out[b][k][l][j], k in [phi(m')/phi(m)]: the O_m-coefficients of in_hi w.r.t. the relative
 * powerful (or decoding) basis; ab_uncoeffs is its inverse (sum_k rel_k * embed(c_k), as an index scatter). */
int ab_coeffs_dev(const ab_ext *e, int64_t *out_lo, const int64_t *in_hi, int64_t batch, int32_t nlimbs, void *stream);
int ab_uncoeffs_dev(const ab_ext *e, int64_t *out_hi, const int64_t *in_lo, int64_t batch, int32_t nlimbs, void *stream);
int64_t ab_ext_rel_dim(const ab_ext *e);       /* phi(m')/phi(m) */

/* ---- cross-limb (RNS) operations used by SymmSHE; PLANAR layout, powerful/decoding basis ---------
 * Gadget decomposition with Lol's product-ring gadget (decompose (a,b) = decompose a ++ decompose b):
 *   base == 0 : TrivGad   -- one digit per limb, the centred lift of that limb's residue
 *   base == b : BaseBGad b -- ceil(log_b q_l) signed digits per limb (divModCent: -b/2 <= r < b/2)
 * and `reduce` of every digit into each of the `out_limbs` leading moduli:
 *   out[b][d][l'][j] = digit_d(in[b][.][j]) mod q_l'      d in [ab_gadget_len], l' in [out_limbs]. */
int32_t ab_gadget_len(const ab_plan *p, int32_t nlimbs, int32_t base);
int ab_decompose_dev(const ab_plan *p, int64_t *out, const int64_t *in, int64_t batch, int32_t nlimbs, int32_t out_limbs, int32_t base, void *stream);
/* Lol `RescaleCyc (a,b) -> b` applied to drop limb `drop` (in Pow/Dec basis):
 *   out_l = (in_l - centredlift(in_drop)) * q_drop^-1  mod q_l    for every kept limb l.
 * `in` has nlimbs limbs, `out` has nlimbs-1 (kept limbs in order).  out may alias in only if drop is last. */
int ab_rescale_drop_dev(const ab_plan *p, int64_t *out, const int64_t *in, int64_t batch, int32_t nlimbs, int32_t drop, void *stream);

/* ---- fused ciphertext-level operations (upstream has no C entry for these; Lol does them in Haskell)
 * Ciphertext polynomials are stored coefficient-major: c[b][i][l][j], i = degree in the secret.
 * All inputs/outputs in the CRT basis, PLANAR layout. */

/* SymmSHE (*) : (c0,c1)*(d0,d1) -> (c0d0, c0d1+c1d0, c1d1), every coefficient times g (mulGCRT).
 * Eval.hs:65-67 `mul_ = pureE (*)`. */
int ab_ct_mul_dev(const ab_plan *p, int64_t *out3, const int64_t *a2, const int64_t *b2, int64_t batch, int32_t nlimbs, void *stream);

/* modSwitch = rescaleCyc on every coefficient of an MSD ciphertext (Eval.hs:130), from the first `from_limbs` moduli to
 * the first `to_limbs` moduli of the plan.  to < from: drop limbs from the end (crtInv of the dropped limbs, centred lift,
 * CRT into the kept limbs, subtract, scale; the result equals dropping them one at a time);
 * to > from: multiply by the new moduli and append zero limbs.  `npoly` polynomials per ciphertext (<= 31).
 * `dec_mask`: bit i set = polynomial i is rounded in the DECODING basis (rescaleDec), clear = in the powerful basis (rescalePow).
 *   AB_RESCALE_POW (0) rounds every polynomial in the powerful basis; AB_RESCALE_LOL_LINEAR (1) is Lol's rescaleLinearMSD
 *   [RECALL]: c0 through rescaleDec, every further coefficient through rescalePow.  Rings without odd primes have Dec = Pow.
 * The caller brings the ciphertext to MSD form first (toMSD: ab_scalar_mul_dev by p^-1 mod q_l) -- see ab_mul_key_switch_dev.
 * `work`: device scratch of ab_mod_switch_work_bytes() bytes. */
enum { AB_RESCALE_POW = 0, AB_RESCALE_LOL_LINEAR = 1 };
int64_t ab_mod_switch_work_bytes(const ab_plan *p, int64_t batch, int32_t npoly, int32_t from_limbs, int32_t to_limbs);
int ab_mod_switch_dev(const ab_plan *p, int64_t *out, const int64_t *in, int64_t batch, int32_t npoly,
                      int32_t from_limbs, int32_t to_limbs, uint32_t dec_mask, void *work, void *stream);

/* keySwitchQuadCirc (Eval.hs:133): (c0,c1,c2) -> (c0,c1) + sum_d reduce(decompose(c2))_d * hint_d.
 * `hint`: [gadget_len][2][nlimbs][n] in the CRT basis, shared by the whole batch. */
int64_t ab_key_switch_work_bytes(const ab_plan *p, int64_t batch, int32_t nlimbs, int32_t base);
int ab_key_switch_quad_dev(const ab_plan *p, int64_t *out2, const int64_t *in3, const int64_t *hint,
                           int64_t batch, int32_t nlimbs, int32_t base, void *work, void *stream);

/* PT2CT `mul_` (PT2CT.hs:172-177) as one call:  modSwitch(l_out) . keySwitchQuad hint . modSwitch(l_hint) . (*)
 * a2, b2: [batch][2][l_in][n] (LSD forms); hint: [gadget_len(l_hint)][2][l_hint][n]; out2: [batch][2][l_out][n] (MSD form).
 * `prescale`: HOST array of l_in residues or NULL.  Lol's (*) yields an LSD ciphertext and modSwitch begins with toMSD, which
 *   multiplies every polynomial by p^-1 mod q_l (ModSwitchCtx_ = (RescaleCyc, ToSDCtx), Eval.hs:124): pass prescale[l] = p^-1 mod q_l
 *   to get exactly  modSwitch . keySwitchQuadCirc . modSwitch . (*)  of Lol; it is folded into the product at no cost.  NULL = 1.
 *   The (k, l) bookkeeping of the result (k = k_a + k_b + 1, l = l_a l_b (-Q_in) mod p) stays with the caller.
 * `dec_mask`: rounding bases of both modSwitch steps, as in ab_mod_switch_dev.
 * Power-of-two rings that fit shared memory (examples/Arithmetic.hs: m' = 512) with TrivGad run as a single
 * kernel, every intermediate on chip; all other cases as the sequence of the entries above. */
int64_t ab_mul_key_switch_work_bytes(const ab_plan *p, int64_t batch, int32_t l_in, int32_t l_hint, int32_t l_out, int32_t base);
int ab_mul_key_switch_dev(const ab_plan *p, int64_t *out2, const int64_t *a2, const int64_t *b2, const int64_t *hint, int64_t batch,
                          int32_t l_in, int32_t l_hint, int32_t l_out, int32_t base, const int64_t *prescale, uint32_t dec_mask,
                          void *work, void *stream);

/* tunnel (Eval.hs:134): ring switch O_r' -> O_s' through E' = O_e'.
 *   c0' = sum_k embed(coeffsDec(c0)_k) * flin_k            (evalLin of the E'-linear function)
 *   (c0'',c1'') = sum_k sum_d reduce(decompose(embed(coeffsPow(c1)_k)))_d * hint_{k,d}
 * `flin`: [rel_dim][nlimbs][n_s] CRT basis of O_s';  `hint`: [rel_dim][gadget_len][2][nlimbs][n_s].
 * Input (c0,c1) over O_r' in the CRT basis; output (c0,c1) over O_s' in the CRT basis.
 * plan_e needs no CRT tables: ab_tunnel_create builds its own CRT plan of e' over the moduli of plan_s (root omega_s'^(s'/e')) and
 * runs every gadget-digit transform in O_e' -- CRT_s'(embed x) is CRT_e'(x) replicated through Lol's embedCRT index. */
typedef struct ab_tunnel ab_tunnel;
int ab_tunnel_create(ab_tunnel **out, const ab_plan *plan_r, const ab_plan *plan_e, const ab_plan *plan_s);
void ab_tunnel_destroy(ab_tunnel *t);
int64_t ab_tunnel_work_bytes(const ab_tunnel *t, int64_t batch, int32_t nlimbs, int32_t base);
int ab_tunnel_dev(const ab_tunnel *t, int64_t *out2_s, const int64_t *in2_r, const int64_t *flin, const int64_t *hint,
                  int64_t batch, int32_t nlimbs, int32_t base, void *work, void *stream);

/* ---- the ciphertext <-> plaintext boundary of SymmSHE (upstream: Haskell over Cyc values [RECALL]) ------------
 * All PLANAR; ciphertexts c[b][i][l][j] in the CRT basis over the first `nlimbs` moduli of the plan. */

/* addPublic / mulPublic cores (Eval.hs:131-132): `pub`[l][j] is the public ring element already brought to the
 * ciphertext's form (embedded into O_m', reduced mod every q_l, times g^k / scaled by l^-1, CRT basis):
 *   add: c[b][0] += pub        mul: c[b][i] *= pub for every i < npoly */
int ab_ct_add_public_dev(const ab_plan *p, int64_t *c, const int64_t *pub, int64_t batch, int32_t npoly, int32_t nlimbs, void *stream);
int ab_ct_mul_public_dev(const ab_plan *p, int64_t *c, const int64_t *pub, int64_t batch, int32_t npoly, int32_t nlimbs, void *stream);

/* `evaluate c s` of Lol's errorTermUnrestricted / decrypt (PT2CT.hs:99, Eval.hs:155-160):
 *   out[b][l][j] = sum_i c[b][i][l][j] * s[l][j]^i     (s = the secret key reduced mod q, CRT basis) */
int ab_ct_phase_dev(const ab_plan *p, int64_t *out, const int64_t *c, const int64_t *s, int64_t batch, int32_t npoly, int32_t nlimbs, void *stream);

/* `reduce . lift` across the RNS product, per coefficient (any basis):  x = centred representative of in[b][.][j]
 * modulo Q = q_0 ... q_(nlimbs-1), in (-Q/2, Q/2);  out[b][j] = x mod pmod in [0, pmod).
 * maxrate (may be NULL): maxrate[b] = max_j |x| / Q as a double -- ALCHEMY's errorRate (Eval.hs:155-160). */
int ab_rns_lift_reduce_dev(const ab_plan *p, int64_t *out, double *maxrate, const int64_t *in, int64_t batch, int32_t nlimbs, int64_t pmod, void *stream);

/* plaintext side of encrypt / addPublic / mulPublic: out[b][l][j] = lift_pmod(in[b][j]) mod q_l, with Lol's
 * `lift` = the representative in [-pmod/2, pmod/2).  `in` holds one coefficient vector (n of the plan) per b. */
int ab_lift_plain_dev(const ab_plan *p, int64_t *out, const int64_t *in, int64_t batch, int32_t nlimbs, int64_t pmod, void *stream);

/* ---- randomness for keys and hints (KeysHints.hs:96,113,115-129; upstream: Haskell MonadRandom + tGaussianDec) --
 * Counter-based (Philox4x32-10): every value is a function of (seed, stream_id, position) only.
 * FOR TESTS AND BENCHMARKS: Philox is a statistical generator with a 64-bit key, not a CSPRNG, and a repeated (seed, stream_id)
 * repeats its output -- callers must give every use its own stream_id (the host mirror folds a per-call nonce into it).  The error
 * coefficients are i.i.d. in the decoding basis, NOT Lol's tGaussianDec (spherical in the canonical embedding): for m that is not a
 * power of two the two distributions differ.  Production keys: draw masks and noise elsewhere and pass them to the entries above.
 *   ab_uniform_dev : out[b][l][j] uniform in [0, q_l)
 *   ab_gaussian_dev: out[b][j] signed integers, |x| = #{k : cdt[k] <= u} for a uniform 64-bit u, with
 *                    cdt[k] = floor(2^64 P(|X| <= k)) (HOST array, non-decreasing, at most 128 entries) */
int ab_uniform_dev(const ab_plan *p, int64_t *out, int64_t batch, int32_t nlimbs, uint64_t seed, uint32_t stream_id, void *stream);
int ab_gaussian_dev(const ab_plan *p, int64_t *out, int64_t batch, const uint64_t *cdt, int32_t cdt_len, uint64_t seed, uint32_t stream_id, void *stream);

/* ---- wire / on-disk format (upstream: none -- keys and hints are in-memory Dynamics, KeysHints.hs:36-39) -------------
 * 136-byte little-endian header (format version 2), then one uint32 per residue in the device layout [batch][npoly][nlimbs][n]
 * (residues are below 2^31: half the bytes of the int64 tensors on NVLink / PCIe / disk). */
typedef struct {
    uint32_t magic;            /* "AB2W"; filled in by ab_wire_pack_dev */
    uint16_t version, kind;    /* kind: 1 ring elements, 2 ciphertexts, 3 key-switch hint, 4 tunnel hint */
    int64_t m, n, batch;       /* cyclotomic index of the polynomials, phi(m), leading dimension */
    int32_t npoly, nlimbs;     /* polynomials per unit (ciphertext: degree + 1; hint row: 2), RNS limbs */
    int32_t enc, k;            /* SymmSHE bookkeeping of ciphertexts: 0 LSD / 1 MSD, power of g */
    int64_t l, pmod;           /* scalar l, plaintext modulus */
    int64_t moduli[8];
    uint64_t checksum;         /* sum_j (j+1) K w_j + sum_i (i+1) v_i mod 2^64 over the sixteen 64-bit header words w_j before this field
                                  (K = 0x9E3779B97F4A7C15) and the payload words v_i: covers every header field and the ORDER of the
                                  payload; filled in by ab_wire_pack_dev, verified by ab_wire_unpack_dev */
} ab_wire_header;
int64_t ab_wire_bytes(const ab_wire_header *h);                    /* header + payload */
/* wire_dev: DEVICE buffer of ab_wire_bytes(h) bytes, 8-byte aligned; h: HOST header (magic/version/checksum are overwritten).
 * Fails with AB_ERR_ARG if any value of data_dev is not a 31-bit residue (negative or >= 2^31); returns after the header is written. */
int ab_wire_pack_dev(void *wire_dev, const ab_wire_header *h, const int64_t *data_dev, void *stream);
/* validates magic, version and that wire_bytes matches the header; synchronous */
int ab_wire_read_header(ab_wire_header *h_out, const void *wire_dev, int64_t wire_bytes);
/* unpacks into data_dev ([batch][npoly][nlimbs][n] int64) and verifies the checksum; returns after the check */
int ab_wire_unpack_dev(int64_t *data_dev, const void *wire_dev, int64_t wire_bytes, void *stream);

/* ---- host-buffer entries: what a `foreign import ccall` from Lol's CPP.hs would bind ------------
 * lol-cpp shape [RECALL]: void tensorCRTRq(hShort_t tupSize, Zq* y, hDim_t totm, PrimeExponent* pe,
 *                                         hShort_t sizeOfPE, Zq** ru, hInt_t* qs)
 * Ours: same data (interleaved tuples in host memory, in place), plus a batch count and a status.
 * `op` selects the Tensor method. */
enum {
    AB_OP_CRT = 0, AB_OP_CRT_INV = 1, AB_OP_L = 2, AB_OP_L_INV = 3,
    AB_OP_MUL_G_POW = 4, AB_OP_MUL_G_DEC = 5, AB_OP_DIV_G_POW = 6, AB_OP_DIV_G_DEC = 7,
    AB_OP_MUL_G_CRT = 8, AB_OP_DIV_G_CRT = 9
};
int ab_tensor_op_host(const ab_plan *p, int op, int64_t *y_host, int64_t batch, int layout);
int ab_mul_host(const ab_plan *p, int64_t *y_host, const int64_t *x_host, int64_t batch, int layout);

/* drop-in names with the upstream argument order [RECALL]; plans are cached per (pe, qs) inside. */
int tensorCRTRq(int16_t tupSize, int64_t *y, int64_t totm, const ab_prime_exp *pe, int16_t sizeOfPE, const int64_t *const *ru, const int64_t *qs);
int tensorCRTInvRq(int16_t tupSize, int64_t *y, int64_t totm, const ab_prime_exp *pe, int16_t sizeOfPE, const int64_t *const *ruinv, const int64_t *mhatInv, const int64_t *qs);
int tensorLRq(int16_t tupSize, int64_t *y, int64_t totm, const ab_prime_exp *pe, int16_t sizeOfPE, const int64_t *qs);
int tensorLInvRq(int16_t tupSize, int64_t *y, int64_t totm, const ab_prime_exp *pe, int16_t sizeOfPE, const int64_t *qs);
int tensorGPowRq(int16_t tupSize, int64_t *y, int64_t totm, const ab_prime_exp *pe, int16_t sizeOfPE, const int64_t *qs);
int tensorGDecRq(int16_t tupSize, int64_t *y, int64_t totm, const ab_prime_exp *pe, int16_t sizeOfPE, const int64_t *qs);
int tensorGInvPowRq(int16_t tupSize, int64_t *y, int64_t totm, const ab_prime_exp *pe, int16_t sizeOfPE, const int64_t *qs);
int tensorGInvDecRq(int16_t tupSize, int64_t *y, int64_t totm, const ab_prime_exp *pe, int16_t sizeOfPE, const int64_t *qs);
int mulRq(int16_t tupSize, int64_t *a, const int64_t *b, int64_t totm, const int64_t *qs);

/* ---- instrumentation --------------------------------------------------------------------------- */
uint64_t ab_kernel_launches(void);      /* kernels launched by this library in this process */

#ifdef __cplusplus
}
#endif
#endif
