/*
 * alchemy_oracle.c -- CPU restatement of the ring-arithmetic hot path behind
 * ALCHEMY's Lol `Tensor` backend.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this file.  The product
 * (alchemy_b200/) never links or calls it.
 *
 * PARITY UNPINNED.  The arithmetic of this path is not in /root/reference: it lives in
 * the un-vendored packages lol / lol-apps / lol-cpp (stack.yaml:55-60, branch
 * `alchemy-args-debruijn-monad`; alchemy.cabal:51-52 `lol >= 0.7`, `lol-apps >= 0.2`),
 * the reference ships no tests or golden vectors (alchemy.cabal:22-123 has no test-suite),
 * and no Haskell toolchain exists here.  This file therefore restates the *published
 * algorithm* (Crockett-Peikert, "Lambda o lambda: Functional Lattice Cryptography",
 * App. C "sparse decompositions"; Lyubashevsky-Peikert-Regev, "A Toolkit for Ring-LWE
 * Cryptography") and is anchored on the reference's call sites:
 *   Crypto/Alchemy/Interpreter/Eval.hs:58-67,95,129-134   (the ops that reach the backend)
 *   Crypto/Alchemy/Interpreter/PT2CT.hs:172-177,224-229   (op sequences)
 *   examples/Common.hs:35                                  (Zq q = ZqBasic q Int64)
 * Every order/representative convention on which bit-exactness against lol-cpp would
 * depend is collected in the "CONVENTIONS" block below.
 * Each function has two forms where it matters: a *definitional* one (straight from the
 * mathematics, O(n^2)) and a *structured* one (the sparse decomposition); tests pin one
 * against the other and against convention-independent invariants.
 *
 * CONVENTIONS (all [RECALL] of upstream Lol; unverifiable here)
 *  C1  residues are int64 in [0,q).
 *  C2  omega_m = g^((q-1)/m), g = smallest generator of Z_q^*.  omega_{m/d} = omega_m^d.
 *  C4  powerful basis of a prime power p^e: tensor index j = j1*(p-1) + j0,
 *      j1 in [p^(e-1)], j0 in [p-1], stands for zeta^(p^(e-1)*j0 + digitrev_{e-1}(j1)).
 *  C3  CRT rows of a prime power: index i = i1*(p-1) + i0 stands for the unit
 *      p*i1 + i0 + 1 of Z_{p^e}^*  (evaluation at omega_{p^e}^(p*i1+i0+1)).
 *      Arbitrary m: Kronecker product over prime powers, primes ascending, first
 *      prime slowest; units combine by the Chinese remainder theorem.
 *  C9  decoding basis: pow = L * dec with L_{p^e} = I_{p^(e-1)} (x) L_p, L_p the lower
 *      triangular all-ones matrix (running sum along j0).
 *  C6  g = prod over ODD primes p | m of (1 - zeta_p).
 *  C5/C7/C8 (tuple layout, gadget digits, rescale rounding) live in she_oracle.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

typedef struct { int32_t p; int32_t e; } ao_pp;

/* ------------------------------------------------------------------ modular basics */
static inline int64_t mulmod(int64_t a, int64_t b, int64_t q) {
    /* lol-cpp style [RECALL]: plain 64-bit multiply and `%`; exact because q < 2^31 */
    if (q < ((int64_t)1 << 31)) return (int64_t)(((uint64_t)a * (uint64_t)b) % (uint64_t)q);
    return (int64_t)(((unsigned __int128)(uint64_t)a * (uint64_t)b) % (uint64_t)q);
}
static inline int64_t addmod(int64_t a, int64_t b, int64_t q) { int64_t s = a + b; return s >= q ? s - q : s; }
static inline int64_t submod(int64_t a, int64_t b, int64_t q) { int64_t s = a - b; return s < 0 ? s + q : s; }

int64_t ao_powmod(int64_t b, int64_t e, int64_t q) {
    int64_t r = 1 % q; b %= q; if (b < 0) b += q;
    while (e > 0) { if (e & 1) r = mulmod(r, b, q); b = mulmod(b, b, q); e >>= 1; }
    return r;
}
int64_t ao_invmod(int64_t a, int64_t q) { return ao_powmod(a, q - 2, q); } /* q prime */

static int64_t ipow(int64_t b, int e) { int64_t r = 1; while (e-- > 0) r *= b; return r; }

/* factor m into prime powers, primes ascending (Lol `ppsFact`); returns count */
int ao_factor(int64_t m, ao_pp *out) {
    int n = 0;
    for (int64_t p = 2; p * p <= m; ++p) {
        if (m % p == 0) { int e = 0; while (m % p == 0) { m /= p; ++e; } out[n].p = (int32_t)p; out[n].e = e; ++n; }
    }
    if (m > 1) { out[n].p = (int32_t)m; out[n].e = 1; ++n; }
    return n;
}
static int64_t pp_value(ao_pp pp) { return ipow(pp.p, pp.e); }
static int64_t pp_totient(ao_pp pp) { return ipow(pp.p, pp.e - 1) * (pp.p - 1); }
int64_t ao_value(int npp, const ao_pp *pps) { int64_t m = 1; for (int i = 0; i < npp; ++i) m *= pp_value(pps[i]); return m; }
int64_t ao_totient(int npp, const ao_pp *pps) { int64_t n = 1; for (int i = 0; i < npp; ++i) n *= pp_totient(pps[i]); return n; }

/* C2: smallest generator of Z_q^* (q prime), scanning 2,3,... (Lol ZqBasic principalRootUnity) */
int64_t ao_smallest_generator(int64_t q) {
    int64_t order = q - 1, fs[64]; int nf = 0; int64_t t = order;
    for (int64_t p = 2; p * p <= t; ++p) if (t % p == 0) { fs[nf++] = p; while (t % p == 0) t /= p; }
    if (t > 1) fs[nf++] = t;
    for (int64_t g = 1; g < q; ++g) {
        int ok = 1;
        for (int i = 0; i < nf && ok; ++i) if (ao_powmod(g, order / fs[i], q) == 1) ok = 0;
        if (ok && (q == 2 || g != 1 || order == 1)) return g;
    }
    return 0;
}
/* omega_m for modulus q; 0 if m does not divide q-1 */
int64_t ao_root_of_unity(int64_t q, int64_t m) {
    if ((q - 1) % m) return 0;
    return ao_powmod(ao_smallest_generator(q), (q - 1) / m, q);
}

/* ------------------------------------------------------------------ index conventions */
/* base-p digit reversal of j over e digits (Lol `digitRev`) */
int64_t ao_digit_rev(int32_t p, int32_t e, int64_t j) {
    int64_t r = 0; for (int i = 0; i < e; ++i) { r = r * p + j % p; j /= p; } return r;
}
/* C4: power of zeta_{p^e} for tensor index j (Lol `indexToPow`) */
static int64_t pp_index_to_pow(ao_pp pp, int64_t j) {
    int64_t jq = j / (pp.p - 1), jr = j % (pp.p - 1);
    return ipow(pp.p, pp.e - 1) * jr + ao_digit_rev(pp.p, pp.e - 1, jq);
}
/* C3: unit of Z_{p^e}^* for CRT index i (Lol `indexToZms`) */
static int64_t pp_index_to_zms(ao_pp pp, int64_t i) {
    int64_t i1 = i / (pp.p - 1), i0 = i % (pp.p - 1);
    return pp.p * i1 + i0 + 1;
}
/* out[j] = exponent k (mod m) with pow-basis element j = zeta_m^k */
void ao_pow_exponents(int npp, const ao_pp *pps, int64_t *out) {
    int64_t m = ao_value(npp, pps), n = ao_totient(npp, pps);
    for (int64_t j = 0; j < n; ++j) {
        int64_t rem = j, k = 0;
        for (int a = npp - 1; a >= 0; --a) {
            int64_t d = pp_totient(pps[a]), ja = rem % d; rem /= d;
            k = (k + (m / pp_value(pps[a])) % m * pp_index_to_pow(pps[a], ja)) % m;
        }
        out[j] = k;
    }
}
/* out[i] = the unit u of Z_m^* with CRT row i = evaluation zeta_m -> omega_m^u.
 * Each axis evaluates zeta_pp -> omega_pp^(z_a) with omega_pp = omega_m^(m/pp) = g^((q-1)/pp)
 * and z_a = zms_a(i_a); since zeta_m = prod_a zeta_pp^(c_a) with c_a = (m/pp_a)^-1 mod pp_a,
 * the row sends zeta_m to omega_m^u with u = z_a (mod pp_a) for every a: the Chinese-remainder
 * combination (Lol `zmsToIndex` reduces the unit mod each prime power, [RECALL]). */
static int64_t inv_mod_small(int64_t a, int64_t n) { /* a^-1 mod n by brute force on small n */
    a %= n; for (int64_t x = 1; x < n; ++x) if ((a * x) % n == 1) return x; return n == 1 ? 0 : -1;
}
void ao_crt_units(int npp, const ao_pp *pps, int64_t *out) {
    int64_t m = ao_value(npp, pps), n = ao_totient(npp, pps), coef[16];
    for (int a = 0; a < npp; ++a) {
        int64_t v = pp_value(pps[a]), co = m / v;
        coef[a] = (co * inv_mod_small(co, v)) % m;      /* = 1 mod pp_a, 0 mod the others */
    }
    for (int64_t i = 0; i < n; ++i) {
        int64_t rem = i, u = 0;
        for (int a = npp - 1; a >= 0; --a) {
            int64_t d = pp_totient(pps[a]), ia = rem % d; rem /= d;
            u = (u + (int64_t)(((__int128)coef[a] * pp_index_to_zms(pps[a], ia)) % m)) % m;
        }
        out[i] = u;
    }
}

/* ------------------------------------------------------------------ definitional CRT */
/* y[i] = sum_j x[j] * omega_m^(unit(i) * powexp(j)).   O(n^2); small cases only. */
void ao_crt_def(int npp, const ao_pp *pps, int64_t q, const int64_t *x, int64_t *y) {
    int64_t m = ao_value(npp, pps), n = ao_totient(npp, pps);
    int64_t w = ao_root_of_unity(q, m);
    int64_t *pe = malloc(sizeof(int64_t) * n), *un = malloc(sizeof(int64_t) * n), *wp = malloc(sizeof(int64_t) * m);
    ao_pow_exponents(npp, pps, pe); ao_crt_units(npp, pps, un);
    wp[0] = 1; for (int64_t k = 1; k < m; ++k) wp[k] = mulmod(wp[k - 1], w, q);
    for (int64_t i = 0; i < n; ++i) {
        int64_t acc = 0;
        for (int64_t j = 0; j < n; ++j)
            acc = addmod(acc, mulmod(x[j], wp[(int64_t)(((__int128)un[i] * pe[j]) % m)], q), q);
        y[i] = acc;
    }
    free(pe); free(un); free(wp);
}

/* ------------------------------------------------------------------ structured CRT */
/* Per prime power (Lol paper App. C):
 *   CRT_{p^e} = (DFT_{p^(e-1)} (x) I_{p-1}) * That * (I_{p^(e-1)} (x) CRT_p)
 *   DFT_{p^k} = (DFT_{p^(k-1)} (x) I_p)     * T    * (I_{p^(k-1)} (x) DFT_p)
 * Realised here as: CRT_p on the fast (j0) axis; twiddle omega_{p^e}^(rev(j1)*(i0+1));
 * then a decimation-in-time radix-p DFT over j1 (input digit-reversed, output natural).
 * An axis lives inside the Kronecker product with `lts` blocks to its left and stride
 * `rts` to its right (lol-cpp's tensor fuser, [RECALL]). */
typedef struct { int64_t q; ao_pp pp; int64_t *wpow; /* omega_{p^e}^k, k in [p^e] */ } axis_ctx;

static void axis_ctx_init(axis_ctx *c, int64_t q, ao_pp pp, int64_t w_pp, int inverse) {
    int64_t v = pp_value(pp);
    c->q = q; c->pp = pp; c->wpow = malloc(sizeof(int64_t) * v);
    if (inverse) w_pp = ao_invmod(w_pp, q);
    c->wpow[0] = 1; for (int64_t k = 1; k < v; ++k) c->wpow[k] = mulmod(c->wpow[k - 1], w_pp, q);
}

/* forward CRT along one axis vector v (length d = phi(p^e), stride s) */
static void axis_crt_fwd(const axis_ctx *c, int64_t *v, int64_t s, int64_t *tmp) {
    const int64_t q = c->q; const int32_t p = c->pp.p, e = c->pp.e;
    const int64_t mp = ipow(p, e - 1), pe = mp * p;
    /* step 1: CRT_p on each block of p-1 (omega_p = omega_pp^(p^(e-1))) */
    if (p > 2) {
        for (int64_t j1 = 0; j1 < mp; ++j1) {
            int64_t *b = v + j1 * (p - 1) * s;
            for (int i0 = 0; i0 < p - 1; ++i0) {
                int64_t acc = 0;
                for (int j0 = 0; j0 < p - 1; ++j0)
                    acc = addmod(acc, mulmod(b[j0 * s], c->wpow[(mp * (((int64_t)(i0 + 1) * j0) % p)) % pe], q), q);
                tmp[i0] = acc;
            }
            for (int i0 = 0; i0 < p - 1; ++i0) b[i0 * s] = tmp[i0];
        }
    }
    if (e == 1) return;
    /* step 2: twiddles omega_{p^e}^(rev_{e-1}(j1) * (i0+1)) */
    for (int64_t j1 = 0; j1 < mp; ++j1) {
        int64_t r = ao_digit_rev(p, e - 1, j1);
        for (int i0 = 0; i0 < p - 1; ++i0) {
            int64_t *x = v + (j1 * (p - 1) + i0) * s;
            *x = mulmod(*x, c->wpow[(r * (i0 + 1)) % pe], q);
        }
    }
    /* step 3: DFT_{p^(e-1)} over j1 for each i0; omega_{p^(e-1)} = omega_pp^p.
     * DIT stage t (h = p^t): for block base b, l in [h]: v_c = x[b + c*h + l] * omega_{p*h}^(c*l);
     * x[b + r*h + l] = sum_c omega_p^(c*r) v_c. */
    for (int i0 = 0; i0 < p - 1; ++i0) {
        int64_t *x = v + i0 * s; const int64_t ss = (p - 1) * s; /* stride over j1 */
        for (int64_t h = 1; h < mp; h *= p) {
            const int64_t tw_step = pe / (p * h);          /* omega_{p*h} = omega_pp^(pe/(p*h)) */
            for (int64_t b = 0; b < mp; b += p * h)
                for (int64_t l = 0; l < h; ++l) {
                    for (int cc = 0; cc < p; ++cc)
                        tmp[cc] = mulmod(x[(b + cc * h + l) * ss], c->wpow[(tw_step * ((cc * l) % (p * h))) % pe], q);
                    for (int r = 0; r < p; ++r) {
                        int64_t acc = 0;
                        for (int cc = 0; cc < p; ++cc)
                            acc = addmod(acc, mulmod(tmp[cc], c->wpow[(mp * ((cc * r) % p)) % pe], q), q);
                        x[(b + r * h + l) * ss] = acc;
                    }
                }
        }
    }
}

/* inverse CRT along one axis: exact inverse of axis_crt_fwd.  ctx holds powers of omega^-1. */
static void axis_crt_inv(const axis_ctx *c, int64_t *v, int64_t s, int64_t *tmp) {
    const int64_t q = c->q; const int32_t p = c->pp.p, e = c->pp.e;
    const int64_t mp = ipow(p, e - 1), pe = mp * p;
    const int64_t pinv = ao_invmod(p % q, q);
    if (e > 1) {
        /* inverse of step 3: run the stages backwards with inverse DFT_p and inverse twiddles */
        for (int i0 = 0; i0 < p - 1; ++i0) {
            int64_t *x = v + i0 * s; const int64_t ss = (p - 1) * s;
            for (int64_t h = mp / p; h >= 1; h /= p) {
                const int64_t tw_step = pe / (p * h);
                for (int64_t b = 0; b < mp; b += p * h)
                    for (int64_t l = 0; l < h; ++l) {
                        for (int cc = 0; cc < p; ++cc) {
                            int64_t acc = 0;
                            for (int r = 0; r < p; ++r)
                                acc = addmod(acc, mulmod(x[(b + r * h + l) * ss], c->wpow[(mp * ((cc * r) % p)) % pe], q), q);
                            tmp[cc] = mulmod(mulmod(acc, pinv, q), c->wpow[(tw_step * ((cc * l) % (p * h))) % pe], q);
                        }
                        for (int cc = 0; cc < p; ++cc) x[(b + cc * h + l) * ss] = tmp[cc];
                    }
            }
        }
        /* inverse of step 2 */
        for (int64_t j1 = 0; j1 < mp; ++j1) {
            int64_t r = ao_digit_rev(p, e - 1, j1);
            for (int i0 = 0; i0 < p - 1; ++i0) {
                int64_t *x = v + (j1 * (p - 1) + i0) * s;
                *x = mulmod(*x, c->wpow[(r * (i0 + 1)) % pe], q);
            }
        }
    }
    /* inverse of step 1: a_j = (1/p) * sum_{i=1}^{p-1} ahat_i * (omega^(-i*j) - omega^(i)) */
    if (p > 2) {
        for (int64_t j1 = 0; j1 < mp; ++j1) {
            int64_t *b = v + j1 * (p - 1) * s;
            for (int j0 = 0; j0 < p - 1; ++j0) {
                int64_t acc = 0;
                for (int i = 1; i <= p - 1; ++i) {
                    /* wpow holds omega^-1 powers: omega^(-i*j) = wpow[mp*(i*j % p)], omega^(i) = wpow[mp*((p-i)%p)] */
                    int64_t t = submod(c->wpow[(mp * (((int64_t)i * j0) % p)) % pe], c->wpow[(mp * ((p - i) % p)) % pe], q);
                    acc = addmod(acc, mulmod(b[(i - 1) * s], t, q), q);
                }
                tmp[j0] = mulmod(acc, pinv, q);
            }
            for (int j0 = 0; j0 < p - 1; ++j0) b[j0 * s] = tmp[j0];
        }
    }
}

/* radix-2 fast path for one axis (p == 2): same stages as axis_crt_fwd with the generic
 * index arithmetic hoisted; this is the loop shape a scalar CPU backend runs (table lookup,
 * 64-bit multiply, `%`), so it is what bench.py times as the CPU baseline. */
static void axis2_fwd(const axis_ctx *c, int64_t *v, int64_t s) {
    const int64_t q = c->q, mp = (int64_t)1 << (c->pp.e - 1), pe = mp * 2;
    if (c->pp.e == 1) return;
    for (int64_t j1 = 0; j1 < mp; ++j1) v[j1 * s] = mulmod(v[j1 * s], c->wpow[ao_digit_rev(2, c->pp.e - 1, j1)], q);
    for (int64_t h = 1; h < mp; h *= 2) {
        const int64_t tw = pe / (2 * h);
        for (int64_t b = 0; b < mp; b += 2 * h)
            for (int64_t l = 0; l < h; ++l) {
                int64_t *x0 = v + (b + l) * s, *x1 = v + (b + h + l) * s;
                int64_t u = *x0, t = mulmod(*x1, c->wpow[tw * l], q);
                *x0 = addmod(u, t, q); *x1 = submod(u, t, q);
            }
    }
}

typedef struct { int npp; ao_pp pps[16]; int64_t q, n; axis_ctx fwd[16], inv[16]; } ao_plan;

ao_plan *ao_plan_create(int npp, const ao_pp *pps, int64_t q) {
    ao_plan *pl = calloc(1, sizeof(ao_plan));
    int64_t m = ao_value(npp, pps), w = ao_root_of_unity(q, m);
    if (!w) { free(pl); return NULL; }
    pl->npp = npp; pl->q = q; pl->n = ao_totient(npp, pps);
    for (int a = 0; a < npp; ++a) {
        pl->pps[a] = pps[a];
        int64_t wa = ao_powmod(w, m / pp_value(pps[a]), q);
        axis_ctx_init(&pl->fwd[a], q, pps[a], wa, 0); axis_ctx_init(&pl->inv[a], q, pps[a], wa, 1);
    }
    return pl;
}
void ao_plan_destroy(ao_plan *pl) {
    if (!pl) return;
    for (int a = 0; a < pl->npp; ++a) { free(pl->fwd[a].wpow); free(pl->inv[a].wpow); }
    free(pl);
}
static void plan_crt(const ao_plan *pl, int64_t *y, int inverse) {
    int64_t lts = pl->n, rts = 1, tmp[4096];
    for (int a = pl->npp - 1; a >= 0; --a) {
        const axis_ctx *c = inverse ? &pl->inv[a] : &pl->fwd[a];
        int64_t d = pp_totient(pl->pps[a]); lts /= d;
        for (int64_t l = 0; l < lts; ++l)
            for (int64_t r = 0; r < rts; ++r) {
                int64_t *v = y + l * d * rts + r;
                if (inverse) axis_crt_inv(c, v, rts, tmp);
                else if (pl->pps[a].p == 2) axis2_fwd(c, v, rts);
                else axis_crt_fwd(c, v, rts, tmp);
            }
        rts *= d;
    }
}
void ao_plan_crt(const ao_plan *pl, int64_t *y) { plan_crt(pl, y, 0); }
void ao_plan_crt_inv(const ao_plan *pl, int64_t *y) { plan_crt(pl, y, 1); }
/* generic-path forward CRT (no radix-2 shortcut); tests pin axis2_fwd against it */
void ao_plan_crt_generic(const ao_plan *pl, int64_t *y) {
    int64_t lts = pl->n, rts = 1, tmp[4096];
    for (int a = pl->npp - 1; a >= 0; --a) {
        int64_t d = pp_totient(pl->pps[a]); lts /= d;
        for (int64_t l = 0; l < lts; ++l) for (int64_t r = 0; r < rts; ++r)
            axis_crt_fwd(&pl->fwd[a], y + l * d * rts + r, rts, tmp);
        rts *= d;
    }
}
/* Tensor.crt / Tensor.crtInv on one limb, in place (lol-cpp tensorCRTRq / tensorCRTInvRq [RECALL]) */
void ao_crt(int npp, const ao_pp *pps, int64_t q, int64_t *y) { ao_plan *pl = ao_plan_create(npp, pps, q); plan_crt(pl, y, 0); ao_plan_destroy(pl); }
void ao_crt_inv(int npp, const ao_pp *pps, int64_t q, int64_t *y) { ao_plan *pl = ao_plan_create(npp, pps, q); plan_crt(pl, y, 1); ao_plan_destroy(pl); }
/* `count` independent n-vectors, contiguous, forward (inverse=0) or inverse CRT, on `threads`
 * POSIX threads.  Used by tests for batches and by bench.py for the CPU baseline. */
typedef struct { const ao_plan *pl; int64_t *y; int64_t lo, hi; int inverse; } many_job;
static void *many_worker(void *arg) {
    many_job *j = (many_job *)arg;
    for (int64_t i = j->lo; i < j->hi; ++i) plan_crt(j->pl, j->y + i * j->pl->n, j->inverse);
    return NULL;
}
void ao_plan_crt_many(const ao_plan *pl, int64_t *y, int64_t count, int inverse, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 1024) threads = 1024;
    pthread_t tid[1024]; many_job jobs[1024];
    for (int t = 0; t < threads; ++t) {
        jobs[t] = (many_job){pl, y, count * t / threads, count * (t + 1) / threads, inverse};
        if (t + 1 < threads) pthread_create(&tid[t], NULL, many_worker, &jobs[t]);
    }
    many_worker(&jobs[threads - 1]);
    for (int t = 0; t + 1 < threads; ++t) pthread_join(tid[t], NULL);
}

/* ------------------------------------------------------------------ L, G (per odd/any prime along j0) */
enum { OP_L, OP_LINV, OP_GPOW, OP_GINVPOW, OP_GDEC, OP_GINVDEC };
static void prime_op(int op, int32_t p, int64_t q, int64_t *v, int64_t s) {
    const int d = p - 1;
    int64_t pinv = ao_invmod(p % q, q);
    switch (op) {
    case OP_L:      /* dec -> pow: running sum */
        for (int j = 1; j < d; ++j) v[j * s] = addmod(v[j * s], v[(j - 1) * s], q);
        break;
    case OP_LINV:   /* pow -> dec: differences */
        for (int j = d - 1; j >= 1; --j) v[j * s] = submod(v[j * s], v[(j - 1) * s], q);
        break;
    case OP_GPOW: { /* b_j = a_j - a_{j-1} + a_{p-2} */
        if (p == 2) break;
        int64_t last = v[(d - 1) * s];
        for (int j = d - 1; j >= 1; --j) v[j * s] = addmod(submod(v[j * s], v[(j - 1) * s], q), last, q);
        v[0] = addmod(v[0], last, q); break; }
    case OP_GINVPOW: { /* t = sum(b)/p; a_j = prefix_j(b) - (j+1) t */
        if (p == 2) break;
        int64_t sum = 0; for (int j = 0; j < d; ++j) sum = addmod(sum, v[j * s], q);
        int64_t t = mulmod(sum, pinv, q), pre = 0, jt = 0;
        for (int j = 0; j < d; ++j) { pre = addmod(pre, v[j * s], q); jt = addmod(jt, t, q); v[j * s] = submod(pre, jt, q); }
        break; }
    case OP_GDEC: { /* b_0 = a_0 + sum(a); b_j = a_j - a_{j-1} */
        if (p == 2) break;
        int64_t sum = 0; for (int j = 0; j < d; ++j) sum = addmod(sum, v[j * s], q);
        for (int j = d - 1; j >= 1; --j) v[j * s] = submod(v[j * s], v[(j - 1) * s], q);
        v[0] = addmod(v[0], sum, q); break; }
    case OP_GINVDEC: { /* c_j = sum_{1<=k<=j} b_k; a_0 = (b_0 - sum_j c_j)/p; a_j = a_0 + c_j */
        if (p == 2) break;
        int64_t c = 0, sc = 0;
        for (int j = 1; j < d; ++j) { c = addmod(c, v[j * s], q); sc = addmod(sc, c, q); }
        int64_t a0 = mulmod(submod(v[0], sc, q), pinv, q); c = 0;
        v[0] = a0;
        for (int j = 1; j < d; ++j) { c = addmod(c, v[j * s], q); v[j * s] = addmod(a0, c, q); }
        break; }
    }
}
static void tensor_prime_op(int op, int npp, const ao_pp *pps, int64_t q, int64_t *y) {
    int64_t n = ao_totient(npp, pps), lts = n, rts = 1;
    for (int a = npp - 1; a >= 0; --a) {
        int64_t d = pp_totient(pps[a]), mp = ipow(pps[a].p, pps[a].e - 1); lts /= d;
        /* I_{p^(e-1)} (x) op_p : the p-1 axis is the fast one inside the prime-power block */
        for (int64_t l = 0; l < lts * mp; ++l)
            for (int64_t r = 0; r < rts; ++r)
                prime_op(op, pps[a].p, q, y + l * (pps[a].p - 1) * rts + r, rts);
        rts *= d;
    }
}
void ao_l(int npp, const ao_pp *pps, int64_t q, int64_t *y)          { tensor_prime_op(OP_L, npp, pps, q, y); }
void ao_l_inv(int npp, const ao_pp *pps, int64_t q, int64_t *y)      { tensor_prime_op(OP_LINV, npp, pps, q, y); }
void ao_mul_g_pow(int npp, const ao_pp *pps, int64_t q, int64_t *y)  { tensor_prime_op(OP_GPOW, npp, pps, q, y); }
void ao_div_g_pow(int npp, const ao_pp *pps, int64_t q, int64_t *y)  { tensor_prime_op(OP_GINVPOW, npp, pps, q, y); }
void ao_mul_g_dec(int npp, const ao_pp *pps, int64_t q, int64_t *y)  { tensor_prime_op(OP_GDEC, npp, pps, q, y); }
void ao_div_g_dec(int npp, const ao_pp *pps, int64_t q, int64_t *y)  { tensor_prime_op(OP_GINVDEC, npp, pps, q, y); }

/* g evaluated at every CRT row: out[i] = prod_{odd p | m} (1 - omega_p^(unit(i) mod p)) */
void ao_g_crt(int npp, const ao_pp *pps, int64_t q, int64_t *out) {
    int64_t m = ao_value(npp, pps), n = ao_totient(npp, pps), w = ao_root_of_unity(q, m);
    int64_t *un = malloc(sizeof(int64_t) * n); ao_crt_units(npp, pps, un);
    for (int64_t i = 0; i < n; ++i) {
        int64_t g = 1;
        for (int a = 0; a < npp; ++a) if (pps[a].p != 2) {
            int64_t wp = ao_powmod(w, m / pps[a].p, q);
            g = mulmod(g, submod(1, ao_powmod(wp, un[i] % pps[a].p, q), q), q);
        }
        out[i] = g;
    }
    free(un);
}
void ao_mul_g_crt(int npp, const ao_pp *pps, int64_t q, int64_t *y) {
    int64_t n = ao_totient(npp, pps); int64_t *g = malloc(sizeof(int64_t) * n); ao_g_crt(npp, pps, q, g);
    for (int64_t i = 0; i < n; ++i) y[i] = mulmod(y[i], g[i], q);
    free(g);
}
void ao_div_g_crt(int npp, const ao_pp *pps, int64_t q, int64_t *y) {
    int64_t n = ao_totient(npp, pps); int64_t *g = malloc(sizeof(int64_t) * n); ao_g_crt(npp, pps, q, g);
    for (int64_t i = 0; i < n; ++i) y[i] = mulmod(y[i], ao_invmod(g[i], q), q);
    free(g);
}

/* ------------------------------------------------------------------ pointwise */
void ao_mul(int64_t n, int64_t q, int64_t *y, const int64_t *x) { for (int64_t i = 0; i < n; ++i) y[i] = mulmod(y[i], x[i], q); }
void ao_add(int64_t n, int64_t q, int64_t *y, const int64_t *x) { for (int64_t i = 0; i < n; ++i) y[i] = addmod(y[i], x[i], q); }
void ao_sub(int64_t n, int64_t q, int64_t *y, const int64_t *x) { for (int64_t i = 0; i < n; ++i) y[i] = submod(y[i], x[i], q); }

/* ------------------------------------------------------------------ extension index maps (m | m') */
/* For m | m' (both as prime-power lists over the primes of m'; exponent 0 allowed in `lo`).
 * Pow/Dec: out[j] (j in [phi(m)]) = index in [phi(m')] of the image of basis element j
 *          (Lol extIndicesPowDec).  With C4 each axis maps index j_a -> the same j_a. */
static int64_t pp_tot0(ao_pp pp) { return pp.e == 0 ? 1 : pp_totient(pp); }
void ao_ext_indices_pow(int npp, const ao_pp *hi, const ao_pp *lo, int64_t *out) {
    int64_t n = 1; for (int a = 0; a < npp; ++a) n *= pp_tot0(lo[a]);
    for (int64_t j = 0; j < n; ++j) {
        int64_t rem = j, idx = 0, mul = 1;
        for (int a = npp - 1; a >= 0; --a) {
            int64_t d = pp_tot0(lo[a]), ja = rem % d; rem /= d;
            idx += ja * mul; mul *= pp_totient(hi[a]);
        }
        out[j] = idx;
    }
}
/* CRT: out[i'] (i' in [phi(m')]) = CRT index i in [phi(m)] with unit(i') = unit(i) mod m
 * (Lol baseIndicesCRT).  Per axis: (i1', i0') -> (i1' mod p^(e-1), i0'). */
void ao_base_indices_crt(int npp, const ao_pp *hi, const ao_pp *lo, int64_t *out) {
    int64_t n = 1; for (int a = 0; a < npp; ++a) n *= pp_totient(hi[a]);
    for (int64_t i = 0; i < n; ++i) {
        int64_t rem = i, idx = 0, mul = 1;
        for (int a = npp - 1; a >= 0; --a) {
            int64_t d = pp_totient(hi[a]), ia = rem % d; rem /= d;
            if (lo[a].e == 0) continue;
            int64_t i1 = ia / (hi[a].p - 1), i0 = ia % (hi[a].p - 1);
            idx += ((i1 % ipow(lo[a].p, lo[a].e - 1)) * (lo[a].p - 1) + i0) * mul; mul *= pp_totient(lo[a]);
        }
        out[i] = idx;
    }
}
/* coeffs (Lol `coeffs`, via extIndicesCoeffs): out[k*phi(m) + j] = index in [phi(m')] of
 * (relative basis element k) (x) (basis element j of O_m); k in [phi(m')/phi(m)].
 * Per axis: j1' = j1 + p^(e-1)*k_a (e>=1), or the whole axis index is k_a (e==0). */
void ao_coeff_indices(int npp, const ao_pp *hi, const ao_pp *lo, int64_t *out) {
    int64_t n = 1, np = 1; for (int a = 0; a < npp; ++a) { n *= pp_tot0(lo[a]); np *= pp_totient(hi[a]); }
    int64_t nk = np / n;
    for (int64_t k = 0; k < nk; ++k)
        for (int64_t j = 0; j < n; ++j) {
            int64_t remj = j, remk = k, idx = 0, mul = 1;
            for (int a = npp - 1; a >= 0; --a) {
                int64_t d = pp_tot0(lo[a]), dk = pp_totient(hi[a]) / d;
                int64_t ja = remj % d, ka = remk % dk; remj /= d; remk /= dk;
                int64_t ia;
                if (lo[a].e == 0) ia = ka;
                else { int64_t j1 = ja / (lo[a].p - 1), j0 = ja % (lo[a].p - 1);
                       ia = (j1 + ipow(lo[a].p, lo[a].e - 1) * ka) * (lo[a].p - 1) + j0; }
                idx += ia * mul; mul *= pp_totient(hi[a]);
            }
            out[k * n + j] = idx;
        }
}

