// ab_mixed_plan.cpp -- host side of the fast mixed-radix CRT: compiles CRT_m / CRT_m^-1 into a short list of
// shared-memory passes (AbMxStep, ab_internal.h) with all twiddles folded into small dense matrices.
//
// Same mathematics and index conventions as ab_plan.cpp / the oracle (Lol paper App. C):
//   CRT_{p^e} = (DFT_{p^(e-1)} (x) I_{p-1}) * That * (I_{p^(e-1)} (x) CRT_p),   CRT_m = Kronecker product over p^e | m.
// What changes is the shape handed to the GPU:
//   * `That` never runs as its own pass: the twiddle of row j1 is folded into a per-j1 copy of the CRT_p matrix,
//     the twiddle of a DFT_p stage into a per-row copy of the DFT_p matrix (a few hundred constants per limb);
//   * a prime-power axis of at most 8 points (3^2) is one dense matrix;
//   * the radix-2 stages of the 2^e axis are grouped three at a time (8 points per thread in registers);
//   * the inverse program is the reversed list with every matrix inverted mod q (Gauss-Jordan), so
//     crtInv . crt = id holds by construction;
//   * matrix constants are stored as c * 2^32 mod q: the kernel accumulates sum_c x_c * M_rc in 64 bits and
//     does ONE Montgomery reduction per output instead of one modular reduction per term;
//   * CRT_p of a prime axis p >= 5 (exponent 1: no twiddles folded in) runs as a PAIR step (kind 2).  With w = omega_p, rows i and
//     p - i of the matrix w^(ij) are conjugate and so are columns j and p - j (the missing column p - 1 counts as a zero input):
//         out_i, out_(p-i) = x_0 + sum_{j=1..h} u_j C_ij  +/-  sum_{j=1..h} v_j S_ij,      h = (p-1)/2,
//         u_j = x_j + x_(p-j),  v_j = x_j - x_(p-j),  C_ij = (w^(ij) + w^(-ij)) / 2,  S_ij = (w^(ij) - w^(-ij)) / 2,
//     i.e. p - 1 multiplications per pair of outputs instead of 2 (p - 1).  The inverse has the same shape:
//         x_j = y_j - y_(p-1),   y_j, y_(p-j) = sum_k U_k C'_jk  -/+  sum_k V_k S'_jk,   U_k, V_k = X_k +/- X_(p-k),  C', S' = C, S / p,
//     and y_0 = (sum_k U_k) / p.  64-bit multiply-accumulate (IMAD.WIDE) is the slowest multiply of the SM -- measured 23 lanes
//     per clock per SM on B200 against 63 for a 32-bit IMAD (tools/micro/pipes.cu) -- and these steps were most of them.
// For 2^30 <= q < 2^31 the 64-bit sum of full unsigned products would overflow: those limbs get a "wide" table whose constants
// are the CENTRED representatives (|c| <= (q-1)/2, stored as int32) and the kernel centres its inputs too, so every signed
// product is at most q^2 / 4 < 2^60 in magnitude, up to seven of them fit a signed 64-bit sum, and one signed Montgomery
// reduction per sum still suffices (one IMAD.WIDE per term, as for the narrow moduli; an earlier version split the inputs into
// 16-bit halves against two constant sets: two IMAD.WIDE per term).
#include <algorithm>
#include <cstring>

#include "ab_internal.h"

using namespace abh;

namespace {

struct Builder {
    int L;
    std::vector<AbMxStep> steps;
    std::vector<std::vector<uint32_t>> tab;   // per limb, uint32 words, in the layout of that limb's mode
    std::vector<int> mode;                    // per limb: 1 narrow, 2 wide
    int32_t off_n = 0, off_w = 0;             // running offsets of the two layouts
    Builder(int L_, const std::vector<int> &mode_) : L(L_), tab(L_), mode(mode_) {}
    void pad4() {
        off_n = (off_n + 3) & ~3;
        off_w = (off_w + 3) & ~3;
        for (int l = 0; l < L; ++l)
            while ((int32_t)tab[l].size() < (mode[l] == 2 ? off_w : off_n)) tab[l].push_back(0);
    }
};

uint32_t magic_for(uint32_t d, uint32_t vmax) {
    if (d == 1) return 0;
    const uint32_t mg = (uint32_t)((((uint64_t)1 << 32) + d - 1) / d);
    for (uint32_t v = 0; v <= vmax; ++v)
        if ((uint32_t)(((uint64_t)v * mg) >> 32) != v / d) throw ab_error{AB_ERR_UNSUPPORTED, "mixed plan: division magic failed"};
    return mg;
}

inline int row_stride(int P) { return (P + 3) & ~3; }   // matrix rows padded to 16 bytes

uint64_t to_mont(uint64_t c, uint64_t q) { return (uint64_t)((((unsigned __int128)c) << 32) % q); }
// table word of a constant: Montgomery form; wide limbs (mode 2) store the centred representative as a two's complement int32
uint32_t tab_word(uint64_t c, uint64_t q, int mode) {
    const uint64_t m = to_mont(c, q);
    if (mode == 2 && m > (q - 1) / 2) return (uint32_t)(int32_t)((int64_t)m - (int64_t)q);
    return (uint32_t)m;
}

// a list of P x P matrices (plain residues) -> Montgomery form (wide limbs: centred), rows padded to 16 bytes, appended to limb l's table
void push_matrices(Builder &B, int l, const std::vector<std::vector<uint64_t>> &mats, int P, uint64_t q) {
    const int PP = row_stride(P);
    for (auto &M : mats)
        for (int r = 0; r < P; ++r)
            for (int c = 0; c < PP; ++c) B.tab[l].push_back(c < P ? tab_word(M[r * P + c], q, B.mode[l]) : 0u);
}

// tables of a pair step (kind 2) for the prime p with w = omega_p: row i = 1..h holds C_i1..C_ih, S_i1..S_ih (forward) or the same
// divided by p (inverse, preceded by the row of y_0: h times 1/p, then zeros); Montgomery form (wide limbs: centred), rows padded to 16 bytes.
void push_pair_tables(Builder &B, int l, int p, uint64_t w, uint64_t q, bool inverse) {
    const int P = p - 1, h = P / 2, PP = row_stride(P);
    const uint64_t inv2 = invmod(2 % q, q), invp = invmod((uint64_t)p % q, q), winv = invmod(w, q);
    auto emit_row = [&](const std::vector<uint64_t> &row) {
        for (int c = 0; c < PP; ++c) B.tab[l].push_back(c < (int)row.size() ? tab_word(row[c], q, B.mode[l]) : 0u);
    };
    if (inverse) {
        std::vector<uint64_t> row(P, 0);
        for (int k = 0; k < h; ++k) row[k] = invp;
        emit_row(row);
    }
    for (int i = 1; i <= h; ++i) {
        std::vector<uint64_t> row(P);
        for (int j = 1; j <= h; ++j) {
            const uint64_t a = powmod(w, (uint64_t)i * j % p, q), b = powmod(winv, (uint64_t)i * j % p, q);
            uint64_t c = mulmod((a + b) % q, inv2, q), sgn = mulmod((a + q - b) % q, inv2, q);
            if (inverse) { c = mulmod(c, invp, q); sgn = mulmod(sgn, invp, q); }
            row[j - 1] = c;
            row[h + j - 1] = sgn;
        }
        emit_row(row);
    }
}

std::vector<uint64_t> invert_matrix(std::vector<uint64_t> a, int P, uint64_t q) {
    std::vector<uint64_t> inv((size_t)P * P, 0);
    for (int i = 0; i < P; ++i) inv[i * P + i] = 1;
    for (int col = 0; col < P; ++col) {
        int piv = -1;
        for (int r = col; r < P; ++r)
            if (a[r * P + col] != 0) { piv = r; break; }
        if (piv < 0) throw ab_error{AB_ERR_UNSUPPORTED, "mixed plan: singular stage matrix"};
        if (piv != col)
            for (int c = 0; c < P; ++c) { std::swap(a[piv * P + c], a[col * P + c]); std::swap(inv[piv * P + c], inv[col * P + c]); }
        const uint64_t d = invmod(a[col * P + col], q);
        for (int c = 0; c < P; ++c) { a[col * P + c] = mulmod(a[col * P + c], d, q); inv[col * P + c] = mulmod(inv[col * P + c], d, q); }
        for (int r = 0; r < P; ++r) {
            if (r == col || a[r * P + col] == 0) continue;
            const uint64_t f = a[r * P + col];
            for (int c = 0; c < P; ++c) {
                a[r * P + c] = (a[r * P + c] + q - mulmod(f, a[col * P + c], q)) % q;
                inv[r * P + c] = (inv[r * P + c] + q - mulmod(f, inv[col * P + c], q)) % q;
            }
        }
    }
    return inv;
}

// vec = 2 when two adjacent columns can share one work item (8-byte shared-memory accesses, one set of
// matrix loads for both): needs an even element stride and both columns in the same matrix row
void finish_step(AbMxStep &s) {
    s.vec = (s.estride % 2 == 0 && (s.row_mode != 2 || s.sdiv % 2 == 0)) ? 2 : 1;
    s.div_magic = magic_for((uint32_t)(s.estride / s.vec), (uint32_t)(s.nwork / s.vec) + 1);
    if (s.row_mode == 1) s.row_magic = magic_for((uint32_t)s.rows, (uint32_t)(s.nwork / std::max(1, s.estride)) + 1);
    if (s.row_mode == 2) s.row_magic = magic_for((uint32_t)s.sdiv, (uint32_t)s.estride);
}

AbMxStep matvec_step(int P, int64_t estride, int64_t n, int row_mode, int rows, int64_t sdiv) {
    AbMxStep s{};
    s.kind = 0; s.span = P; s.g = 0; s.estride = (int32_t)estride; s.nwork = (int32_t)(n / P);
    s.row_mode = rows > 1 ? row_mode : 0; s.rows = rows; s.sdiv = (int32_t)sdiv;
    finish_step(s);
    return s;
}

// group sizes for `cnt` radix-2 stages, at most 3 per group, as even as possible
std::vector<int> radix2_groups(int cnt) {
    std::vector<int> g;
    int ng = (cnt + 2) / 3;
    for (int i = 0; i < ng; ++i) g.push_back(cnt / ng + (i < cnt % ng ? 1 : 0));
    return g;
}

}   // namespace

bool ab_mx_span_supported(int P) { return P == 2 || P == 3 || P == 4 || P == 5 || P == 6 || P == 7 || P == 10 || P == 11 || P == 12; }

void ab_build_mixed(ab_plan *pl) {
    const int L = pl->L;
    const int64_t n = pl->n;
    pl->mx_ok.assign(L, 0);
    if (!pl->has_crt || n < 2 || n > 57344) return;          // the transform must fit shared memory as uint32
    // structure must be supported by the kernel's instantiations
    for (auto &ax : pl->axes) {
        if (ax.p == 2) continue;
        const bool dense = ax.e >= 2 && ax.phi <= 8;
        if (dense ? !ab_mx_span_supported((int)ax.phi) : !(ab_mx_span_supported(ax.p - 1) && (ax.e == 1 || ab_mx_span_supported(ax.p)))) return;
    }
    std::vector<int> mode(L);
    for (int l = 0; l < L; ++l) mode[l] = pl->q[l] < (1ll << 30) ? 1 : 2;
    Builder F(L, mode), I(L, mode);
    struct Pending {
        AbMxStep step;
        std::vector<std::vector<std::vector<uint64_t>>> mats;   // forward record, for inversion
        int sym_p = 0;                                             // kind 2: the prime
        std::vector<uint64_t> sym_w;                               // kind 2: omega_p per limb
    };
    static const bool no_sym = getenv("AB_MX_NOSYM") && atoi(getenv("AB_MX_NOSYM")) != 0;      // A/B measurements only
    std::vector<Pending> recs;

    for (int ai = (int)pl->axes.size() - 1; ai >= 0; --ai) {
        const AbAxis &ax = pl->axes[ai];
        const int p = ax.p, e = ax.e;
        const int64_t mp = ax.mp, pp = ax.pp, rts = ax.rts, phi = ax.phi;
        if (p == 2) {
            if (mp == 1) continue;
            int64_t h0 = 1;
            for (int g : radix2_groups(e - 1)) {
                Pending rec;
                AbMxStep &s = rec.step;
                s = AbMxStep{};
                s.kind = 1; s.span = 1 << g; s.g = g; s.estride = (int32_t)(h0 * rts); s.nwork = (int32_t)(n >> g);
                s.row_mode = 2; s.rows = (int32_t)h0; s.sdiv = (int32_t)rts;
                s.flags = (h0 == 1) ? 1 : 0;
                finish_step(s);
                rec.mats.resize(L);
                for (int l = 0; l < L; ++l) {
                    // one pseudo-matrix per stage t: the twiddles psi_{4h}^(2 lt + 1), h = h0 2^t, lt in [h)
                    const uint64_t q = pl->q[l], wpp = powmod(pl->root[l], pl->m / pp, q);
                    for (int t = 0; t < g; ++t) {
                        const int64_t h = h0 << t;
                        const uint64_t base = powmod(wpp, pp / (4 * h), q);
                        std::vector<uint64_t> tw(h);
                        for (int64_t lt = 0; lt < h; ++lt) tw[lt] = powmod(base, 2 * lt + 1, q);
                        rec.mats[l].push_back(tw);
                    }
                }
                recs.push_back(rec);
                h0 <<= g;
            }
            continue;
        }
        const uint64_t P1 = p - 1;
        if (e >= 2 && phi <= 8) {
            Pending rec;
            rec.step = matvec_step((int)phi, rts, n, 0, 1, 1);
            rec.mats.resize(L);
            for (int l = 0; l < L; ++l) {
                const uint64_t q = pl->q[l], wpp = powmod(pl->root[l], pl->m / pp, q);
                std::vector<uint64_t> M((size_t)phi * phi);
                for (int64_t i = 0; i < phi; ++i)
                    for (int64_t j = 0; j < phi; ++j) {
                        const uint64_t unit = (uint64_t)p * (i / P1) + i % P1 + 1;
                        const uint64_t ex = (uint64_t)mp * (j % P1) + (uint64_t)digit_rev(p, e - 1, j / P1);
                        M[i * phi + j] = powmod(wpp, unit * ex % pp, q);
                    }
                rec.mats[l].push_back(M);
            }
            recs.push_back(rec);
            continue;
        }
        {   // CRT_p on the p-1 axis, row twiddles folded in: one matrix per j1
            Pending rec;
            rec.step = matvec_step(p - 1, rts, n, 1, (int)mp, 1);
            rec.mats.resize(L);
            for (int l = 0; l < L; ++l) {
                const uint64_t q = pl->q[l], wpp = powmod(pl->root[l], pl->m / pp, q), wp = powmod(wpp, mp, q);
                for (int64_t j1 = 0; j1 < mp; ++j1) {
                    std::vector<uint64_t> M((size_t)P1 * P1);
                    const uint64_t rv = (uint64_t)digit_rev(p, e - 1, j1);
                    for (uint64_t r = 0; r < P1; ++r)
                        for (uint64_t c = 0; c < P1; ++c)
                            M[r * P1 + c] = mulmod(powmod(wp, (r + 1) * c % p, q), powmod(wpp, rv * (r + 1) % pp, q), q);
                    rec.mats[l].push_back(M);
                }
                rec.sym_w.push_back(wp);
            }
            if (e == 1 && p >= 5 && !no_sym) { rec.step.kind = 2; rec.sym_p = p; }
            recs.push_back(rec);
        }
        const int64_t ss = (p - 1) * rts;
        for (int64_t h = 1; h < mp; h *= p) {   // DFT_p stages over j1, input twiddles folded in: one matrix per row
            Pending rec;
            rec.step = matvec_step(p, h * ss, n, 2, (int)h, ss);
            rec.mats.resize(L);
            for (int l = 0; l < L; ++l) {
                const uint64_t q = pl->q[l], wpp = powmod(pl->root[l], pl->m / pp, q), wp = powmod(wpp, mp, q);
                const uint64_t wph = powmod(wpp, pp / (p * h), q);
                for (int64_t row = 0; row < h; ++row) {
                    std::vector<uint64_t> M((size_t)p * p);
                    for (int r = 0; r < p; ++r)
                        for (int c = 0; c < p; ++c)
                            M[r * p + c] = mulmod(powmod(wp, (uint64_t)c * r % p, q), powmod(wph, (uint64_t)c * row % (p * h), q), q);
                    rec.mats[l].push_back(M);
                }
            }
            recs.push_back(rec);
        }
    }
    if ((int)recs.size() > AB_MX_MAX_STEPS) return;
    // subset-sum tables of the first forward step for binary gadget digits (see ab_internal.h): only when that step is one
    // uniform matrix over contiguous items
    std::vector<std::vector<uint32_t>> bits_lut;
    int bits_p = 0, bits_h = 0;
    if (!recs.empty() && (recs[0].step.kind == 0 || recs[0].step.kind == 2) && recs[0].step.estride == 1 && recs[0].step.row_mode == 0 && recs[0].step.span <= 12 &&
        recs[0].mats[0].size() == 1) {
        bits_p = recs[0].step.span; bits_h = (bits_p + 1) / 2;
        bits_lut.resize(L);
        for (int l = 0; l < L; ++l) {
            const uint64_t q = pl->q[l];
            const auto &M = recs[0].mats[l][0];
            auto &T = bits_lut[l];
            T.assign((size_t)2 * bits_p * (1u << bits_h), 0);
            for (int half = 0; half < 2; ++half)
                for (int r = 0; r < bits_p; ++r)
                    for (uint32_t sub = 0; sub < (1u << bits_h); ++sub) {
                        uint64_t acc = 0;
                        for (int b = 0; b < bits_h; ++b) {
                            const int c = half * bits_h + b;
                            if (((sub >> b) & 1) && c < bits_p) acc = (acc + M[(size_t)r * bits_p + c]) % q;
                        }
                        T[((size_t)half * bits_p + r) * (1u << bits_h) + sub] = (uint32_t)((q - acc) % q);
                    }
        }
    }

    auto emit = [&](Builder &B, const Pending &rec, bool inverse) {
        AbMxStep s = rec.step;
        B.pad4();
        s.tab_off = B.off_n;
        s.tab_off_w = B.off_w;
        for (int l = 0; l < L; ++l) {
            const uint64_t q = pl->q[l];
            if (s.kind == 2) push_pair_tables(B, l, rec.sym_p, rec.sym_w[l], q, inverse);
            else if (s.kind == 0) {
                if (!inverse) push_matrices(B, l, rec.mats[l], s.span, q);
                else {
                    std::vector<std::vector<uint64_t>> inv;
                    for (auto &M : rec.mats[l]) inv.push_back(invert_matrix(M, s.span, q));
                    push_matrices(B, l, inv, s.span, q);
                }
            } else {
                // Shoup pairs; stage t at word offset 2 * h0 * (2^t - 1); inverse: w^-1, and for the group that
                // ends the axis the scale 2^-(e-1) multiplied into stage 0 plus a trailing (scale, scale') pair
                uint64_t scale = 1;
                if (inverse && (s.flags & 1)) {
                    int64_t mp = 1;
                    for (auto &ax : pl->axes) if (ax.p == 2) mp = ax.mp;
                    scale = invmod((uint64_t)mp % q, q);
                }
                for (int t = 0; t < s.g; ++t)
                    for (uint64_t w : rec.mats[l][t]) {
                        uint64_t v = inverse ? invmod(w, q) : w;
                        if (inverse && t == 0) v = mulmod(v, scale, q);
                        B.tab[l].push_back((uint32_t)v);
                        B.tab[l].push_back(shoup(v, q));
                    }
                B.tab[l].push_back((uint32_t)scale);
                B.tab[l].push_back(shoup(scale, q));
            }
        }
        if (s.kind == 2) {
            const int words = (s.span / 2 + (inverse ? 1 : 0)) * row_stride(s.span);      // h (+ 1: the y_0 row) rows of 2 h constants
            B.off_n += words;
            B.off_w += words;
        } else if (s.kind == 0) {
            const int words = s.rows * s.span * row_stride(s.span);
            B.off_n += words;
            B.off_w += words;
        } else {
            const int words = 2 * (s.rows * (s.span - 1) + 1);
            B.off_n += words;
            B.off_w += words;
        }
        B.steps.push_back(s);
    };
    for (auto &rec : recs) emit(F, rec, false);
    for (int i = (int)recs.size() - 1; i >= 0; --i) emit(I, recs[i], true);
    F.pad4();
    I.pad4();
    for (int l = 0; l < L; ++l) {   // the analytic offsets and the pushed tables must agree
        AB_REQUIRE((int32_t)F.tab[l].size() == (mode[l] == 2 ? F.off_w : F.off_n), AB_ERR_UNSUPPORTED, "mixed plan: table layout mismatch");
        AB_REQUIRE((int32_t)I.tab[l].size() == (mode[l] == 2 ? I.off_w : I.off_n), AB_ERR_UNSUPPORTED, "mixed plan: table layout mismatch");
    }
    if ((size_t)std::max(F.off_w, I.off_w) * 4 > 64 * 1024) return;   // tables ride in shared memory next to the data

    pl->mx_fwd = F.steps;
    pl->mx_inv = I.steps;
    pl->mx_words_fwd[0] = F.off_n; pl->mx_words_fwd[1] = F.off_w;
    pl->mx_words_inv[0] = I.off_n; pl->mx_words_inv[1] = I.off_w;
    pl->mx_bits_p = bits_p; pl->mx_bits_h = bits_h;
    pl->mx_bits_lut.assign(L, nullptr);
    for (int l = 0; l < L && bits_p; ++l) {
        uint32_t *dl = nullptr;
        AB_CHECK_CUDA(cudaMalloc((void **)&dl, bits_lut[l].size() * 4));
        pl->owned.push_back(dl);
        AB_CHECK_CUDA(cudaMemcpy(dl, bits_lut[l].data(), bits_lut[l].size() * 4, cudaMemcpyHostToDevice));
        pl->mx_bits_lut[l] = dl;
    }
    pl->mx_tab_fwd.assign(L, nullptr);
    pl->mx_tab_inv.assign(L, nullptr);
    for (int l = 0; l < L; ++l) {
        uint32_t *df = nullptr, *di = nullptr;
        AB_CHECK_CUDA(cudaMalloc((void **)&df, std::max<size_t>(16, F.tab[l].size() * 4)));
        pl->owned.push_back(df);
        AB_CHECK_CUDA(cudaMalloc((void **)&di, std::max<size_t>(16, I.tab[l].size() * 4)));
        pl->owned.push_back(di);
        AB_CHECK_CUDA(cudaMemcpy(df, F.tab[l].data(), F.tab[l].size() * 4, cudaMemcpyHostToDevice));
        AB_CHECK_CUDA(cudaMemcpy(di, I.tab[l].data(), I.tab[l].size() * 4, cudaMemcpyHostToDevice));
        pl->mx_tab_fwd[l] = df;
        pl->mx_tab_inv[l] = di;
        // accumulator bounds of the kernel.  narrow: P (q-1)^2 + (2^32 - 1) q < 2^64 and the reduced value below 4q;
        // wide: at most T = ceil(P / 2) (P > 7), P (dense) or P / 2 (pair step) signed products of centred factors per sum:
        // T ((q-1)/2)^2 < 2^63, and the Montgomery value hi(T) - hi(m q) + 2q (mx_redc_s) must stay inside (0, min(3q, 2^32))
        const unsigned __int128 q = (unsigned __int128)pl->q[l];
        bool ok = true;
        if (mode[l] == 1)
            for (auto &s : F.steps)
                if (s.kind == 0 || s.kind == 2) {      // kind 2: span / 2 terms with a factor below 2 q each
                    ok = ok && ((unsigned __int128)s.span * (q - 1) * (q - 1) + (((unsigned __int128)1 << 32) - 1) * q < ((unsigned __int128)1 << 64));
                    ok = ok && ((unsigned __int128)s.span * q < ((unsigned __int128)3 << 32));
                }
        if (mode[l] == 2)
            for (auto &s : F.steps)
                if (s.kind == 0 || s.kind == 2) {
                    const unsigned __int128 T = s.kind == 2 ? s.span / 2 : (s.span > 7 ? (s.span + 1) / 2 : s.span), hq = (q - 1) / 2;
                    ok = ok && (T * hq * hq < ((unsigned __int128)1 << 63));
                    ok = ok && (((T * hq * hq) >> 32) + 2 * q + 2 < ((unsigned __int128)1 << 32));      // hi(T) - hi(m q) + 2q stays in (0, 2^32)
                    ok = ok && (((T * hq * hq) >> 32) + 2 < q);                                          // ... and below 3q: two conditional subtractions
                }
        pl->mx_ok[l] = ok ? (uint8_t)mode[l] : 0;
    }
}
