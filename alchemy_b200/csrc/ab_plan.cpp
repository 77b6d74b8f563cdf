// ab_plan.cpp -- host side of the backend: number theory, index conventions, table construction.
//
// This is the B200 backend's own statement of the conventions Lol's `Tensor` class fixes
// (SURVEY.md section 8c, C1-C9).  The reference never spells them out: they live in the absent
// `lol` package, reached from Crypto/Alchemy/Interpreter/Eval.hs:58-67,129-134.  [RECALL] items:
//   - omega_m = g^((q-1)/m), g the smallest generator           (Lol ZqBasic principalRootUnity)
//   - powerful-basis index j = j1*(p-1)+j0  <->  zeta^(p^(e-1)*j0 + digitrev(j1))   (indexToPow)
//   - CRT row i = i1*(p-1)+i0  <->  unit p*i1+i0+1                                   (indexToZms)
//   - Kronecker order: primes ascending, first slowest
//   - CRT_{p^e} = (DFT_{p^(e-1)} (x) I) * That * (I (x) CRT_p)  (Lol paper, App. C)
#include <algorithm>
#include <cstring>
#include <map>
#include <mutex>

#include "ab_internal.h"

// ------------------------------------------------------------------------------------ errors
static thread_local std::string g_last_error;
void ab_set_error(const std::string &msg) { g_last_error = msg; }
extern "C" const char *ab_last_error(void) { return g_last_error.c_str(); }
extern "C" const char *ab_version(void) { return "alchemy_b200 0.1 (sm_100a)"; }
extern "C" int ab_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// ------------------------------------------------------------------------------------ number theory
namespace abh {
uint64_t mulmod(uint64_t a, uint64_t b, uint64_t q) { return (uint64_t)(((unsigned __int128)a * b) % q); }
uint64_t powmod(uint64_t b, uint64_t e, uint64_t q) {
    uint64_t r = 1 % q;
    b %= q;
    while (e) {
        if (e & 1) r = mulmod(r, b, q);
        b = mulmod(b, b, q);
        e >>= 1;
    }
    return r;
}
uint64_t invmod(uint64_t a, uint64_t q) { return powmod(a % q, q - 2, q); }
bool is_prime(uint64_t q) {
    if (q < 2) return false;
    for (uint64_t d = 2; d * d <= q; ++d)
        if (q % d == 0) return false;
    return true;
}
uint64_t smallest_generator(uint64_t q) {
    std::vector<uint64_t> fs;
    uint64_t t = q - 1;
    for (uint64_t d = 2; d * d <= t; ++d)
        if (t % d == 0) {
            fs.push_back(d);
            while (t % d == 0) t /= d;
        }
    if (t > 1) fs.push_back(t);
    for (uint64_t g = 1; g < q; ++g) {
        bool ok = true;
        for (uint64_t f : fs)
            if (powmod(g, (q - 1) / f, q) == 1) {
                ok = false;
                break;
            }
        if (ok && (g != 1 || q == 2)) return g;
    }
    return 0;
}
int64_t ipow(int64_t b, int e) {
    int64_t r = 1;
    while (e-- > 0) r *= b;
    return r;
}
int64_t digit_rev(int p, int e, int64_t j) {
    int64_t r = 0;
    for (int i = 0; i < e; ++i) {
        r = r * p + j % p;
        j /= p;
    }
    return r;
}
}   // namespace abh
using namespace abh;

// ------------------------------------------------------------------------------------ helpers
template <class T>
static T *dev_upload(ab_plan *pl, const std::vector<T> &v) {
    T *d = nullptr;
    AB_CHECK_CUDA(cudaMalloc((void **)&d, std::max<size_t>(1, v.size()) * sizeof(T)));
    pl->owned.push_back(d);
    if (!v.empty()) AB_CHECK_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return d;
}
static uint2_ sp(uint64_t w, uint64_t q) { return uint2_{(uint32_t)w, shoup(w, q)}; }

static std::vector<AbAxis> make_axes(const ab_prime_exp *pe, int32_t npe) {
    std::map<int32_t, int32_t> f;
    for (int i = 0; i < npe; ++i) {
        AB_REQUIRE(pe[i].prime >= 2 && pe[i].exponent >= 1, AB_ERR_ARG, "bad prime power");
        AB_REQUIRE(is_prime(pe[i].prime), AB_ERR_ARG, "prime-power base is not prime");
        f[pe[i].prime] += pe[i].exponent;
    }
    std::vector<AbAxis> ax;
    for (auto &kv : f) {
        AbAxis a{};
        a.p = kv.first;
        a.e = kv.second;
        a.mp = ipow(a.p, a.e - 1);
        a.pp = a.mp * a.p;
        a.phi = a.mp * (a.p - 1);
        ax.push_back(a);
    }
    AB_REQUIRE((int)ax.size() <= AB_MAX_AXES, AB_ERR_ARG, "too many distinct primes");
    int64_t n = 1;
    for (auto &a : ax) n *= a.phi;
    int64_t lts = n, rts = 1;
    for (int i = (int)ax.size() - 1; i >= 0; --i) {
        lts /= ax[i].phi;
        ax[i].lts = lts;
        ax[i].rts = rts;
        rts *= ax[i].phi;
    }
    return ax;
}

// unit of Z_m^* evaluated by CRT row i (Chinese-remainder combination of the per-axis units)
static std::vector<int64_t> crt_units(const ab_plan *pl) {
    std::vector<int64_t> coef;
    for (auto &a : pl->axes) {
        int64_t co = pl->m / a.pp, inv = 0;
        for (int64_t x = 1; x < a.pp; ++x)
            if ((co % a.pp) * x % a.pp == 1) {
                inv = x;
                break;
            }
        coef.push_back((int64_t)(((__int128)co * inv) % pl->m));
    }
    std::vector<int64_t> u(pl->n);
    for (int64_t i = 0; i < pl->n; ++i) {
        int64_t rem = i, acc = 0;
        for (int a = (int)pl->axes.size() - 1; a >= 0; --a) {
            const AbAxis &ax = pl->axes[a];
            int64_t ia = rem % ax.phi;
            rem /= ax.phi;
            int64_t z = ax.p * (ia / (ax.p - 1)) + ia % (ax.p - 1) + 1;
            acc = (acc + (int64_t)(((__int128)coef[a] * z) % pl->m)) % pl->m;
        }
        u[i] = acc;
    }
    return u;
}

// ------------------------------------------------------------------------------------ pow2 tables
// Merged negacyclic twiddles (DESIGN.md "pow2 kernel"): the index-2^e CRT with Lol's conventions is
//   y[i] = sum_a x[a] psi^(bitrev(a) * (2i+1)),   psi = omega_{2n},
// a decimation-in-time network whose stage s (distance 2^s) multiplies the upper input by
//   W_s[l] = psi^((2l+1) * n / 2^(s+1)),   l = a mod 2^s.      Table order: tw[2^s + l].
static void build_pow2_tables(ab_plan *pl, int limb, AbLimbDev &ld) {
    const uint64_t q = pl->q[limb], n = pl->n;
    const uint64_t psi = pl->root[limb], psi_inv = invmod(psi, q);
    std::vector<uint2_> fwd(n), inv(n);
    fwd[0] = inv[0] = uint2_{0, 0};
    const uint64_t ninv = invmod(n % q, q);
    for (uint64_t h = 1; h < n; h <<= 1) {
        const uint64_t step = n / (2 * h);
        const uint64_t base = powmod(psi, step, q), base2 = mulmod(base, base, q);
        const uint64_t ibase = powmod(psi_inv, step, q), ibase2 = mulmod(ibase, ibase, q);
        uint64_t w = base, wi = ibase;
        for (uint64_t l = 0; l < h; ++l) {
            fwd[h + l] = sp(w, q);
            // the last inverse stage (h == 1) also carries n^-1
            inv[h + l] = sp(h == 1 ? mulmod(wi, ninv, q) : wi, q);
            w = mulmod(w, base2, q);
            wi = mulmod(wi, ibase2, q);
        }
    }
    // window-packed copies (ab_internal.h) for the kernel geometry chosen for this size
    const int logn = pl->logn, loge = ab_pow2_loge(logn), nw = ab_pow2_nw(logn, loge);
    if (logn >= 5 && logn <= 15) {
        const int64_t total = ab_pow2_base(logn, loge, nw);
        std::vector<uint2_> pf(total, uint2_{0, 0}), pi(total, uint2_{0, 0});
        for (int w = 0; w < nw; ++w) {
            const int S = ab_pow2_s(logn, loge, w), lb0 = ab_pow2_lb0(logn, loge, w), koff = ab_pow2_koff(logn, loge, w);
            const int64_t base = ab_pow2_base(logn, loge, w);
            for (int lb = lb0; lb < loge; ++lb)
                for (int ip = 0; ip < (1 << lb); ++ip)
                    for (int64_t low = 0; low < (1ll << S); ++low) {
                        const int k = (1 << lb) + ip - koff;
                        const int64_t src = (1ll << (S + lb)) + low + ((int64_t)ip << S);
                        const int64_t dst = base + (((int64_t)(k >> 1) << S) + low) * 2 + (k & 1);
                        pf[dst] = fwd[src];
                        pi[dst] = inv[src];
                    }
        }
        ld.tw_fwd = dev_upload(pl, pf);
        ld.tw_inv = dev_upload(pl, pi);
        for (int k = 0; k < 32; ++k) {
            ld.tw0_fwd[k] = (k < (1 << loge) && (uint64_t)k < n) ? fwd[k] : uint2_{0, 0};
            ld.tw0_inv[k] = (k < (1 << loge) && (uint64_t)k < n) ? inv[k] : uint2_{0, 0};
        }
    }
    ld.ninv = (uint32_t)ninv;
    ld.ninv_s = shoup(ninv, q);
    // unpacked copies (tw[2^s + l]) for the shared-memory transforms of the fused mul + key-switch kernel
    ld.tw_fwd_flat = dev_upload(pl, fwd);
    ld.tw_inv_flat = dev_upload(pl, inv);
}

// ------------------------------------------------------------------------------------ generic program
struct ProgBuilder {
    std::vector<AbGenStep> steps;
    std::vector<std::vector<uint2_>> tab;   // per limb
    explicit ProgBuilder(int L) : tab(L) {}
    int32_t off() const { return (int32_t)tab[0].size(); }
};

static void build_generic(ab_plan *pl) {
    const int L = pl->L;
    ProgBuilder F(L), I(L);
    const int64_t n = pl->n;
    // forward: axes last to first (any order is valid: Kronecker factors commute)
    for (int ai = (int)pl->axes.size() - 1; ai >= 0; --ai) {
        const AbAxis &ax = pl->axes[ai];
        const int p = ax.p, e = ax.e;
        const int64_t mp = ax.mp, pp = ax.pp, rts = ax.rts;
        if (p > 2) {
            AbGenStep s{};
            s.kind = 0; s.P = p - 1; s.estride = (int32_t)rts; s.h = (int32_t)mp; s.sdiv = 1;
            s.mode = (e > 1) ? 2 : 0; s.nwork = (int32_t)(n / (p - 1));
            s.mat_off = F.off();
            for (int l = 0; l < L; ++l) {
                const uint64_t q = pl->q[l], wpp = powmod(pl->root[l], pl->m / pp, q), wp = powmod(wpp, mp, q);
                for (int i0 = 0; i0 < p - 1; ++i0)
                    for (int j0 = 0; j0 < p - 1; ++j0) F.tab[l].push_back(sp(powmod(wp, (uint64_t)(i0 + 1) * j0 % p, q), q));
            }
            s.sc_off = F.off();
            if (e > 1)
                for (int l = 0; l < L; ++l) {
                    const uint64_t q = pl->q[l], wpp = powmod(pl->root[l], pl->m / pp, q);
                    for (int64_t j1 = 0; j1 < mp; ++j1)
                        for (int i0 = 0; i0 < p - 1; ++i0)
                            F.tab[l].push_back(sp(powmod(wpp, (uint64_t)digit_rev(p, e - 1, j1) * (i0 + 1) % pp, q), q));
                }
            F.steps.push_back(s);
        }
        // DFT_{p^(e-1)} over j1 (stride ss = (p-1)*rts), decimation in time, h = 1, p, p^2, ...
        const int64_t ss = (p - 1) * rts;
        for (int64_t h = 1; h < mp; h *= p) {
            AbGenStep s{};
            s.P = p; s.estride = (int32_t)(h * ss); s.h = (int32_t)h; s.sdiv = (int32_t)ss;
            s.nwork = (int32_t)(n / p);
            if (p == 2) {
                s.kind = 1; s.mode = 1; s.mat_off = 0; s.sc_off = F.off();
                for (int l = 0; l < L; ++l) {
                    const uint64_t q = pl->q[l], wpp = powmod(pl->root[l], pl->m / pp, q);
                    const uint64_t base = powmod(wpp, pp / (4 * h), q);      // psi_{4h}
                    for (int64_t r = 0; r < h; ++r) F.tab[l].push_back(sp(powmod(base, 2 * r + 1, q), q));
                }
            } else {
                s.kind = 2; s.mode = 1; s.mat_off = F.off();
                for (int l = 0; l < L; ++l) {
                    const uint64_t q = pl->q[l], wpp = powmod(pl->root[l], pl->m / pp, q), wp = powmod(wpp, mp, q);
                    for (int r = 0; r < p; ++r)
                        for (int c = 0; c < p; ++c) F.tab[l].push_back(sp(powmod(wp, (uint64_t)c * r % p, q), q));
                }
                s.sc_off = F.off();
                for (int l = 0; l < L; ++l) {
                    const uint64_t q = pl->q[l], wpp = powmod(pl->root[l], pl->m / pp, q);
                    const uint64_t wph = powmod(wpp, pp / (p * h), q);       // omega_{p*h}
                    for (int64_t r = 0; r < h; ++r)
                        for (int c = 0; c < p; ++c) F.tab[l].push_back(sp(powmod(wph, (uint64_t)c * r % (p * h), q), q));
                }
            }
            F.steps.push_back(s);
        }
    }
    // inverse: per axis (first to last), DFT^-1 stages h = mp/p ... 1, then CRT_p^-1
    for (int ai = 0; ai < (int)pl->axes.size(); ++ai) {
        const AbAxis &ax = pl->axes[ai];
        const int p = ax.p, e = ax.e;
        const int64_t mp = ax.mp, pp = ax.pp, rts = ax.rts;
        const int64_t ss = (p - 1) * rts;
        for (int64_t h = mp / p; h >= 1; h /= p) {
            AbGenStep s{};
            s.P = p; s.estride = (int32_t)(h * ss); s.h = (int32_t)h; s.sdiv = (int32_t)ss;
            s.nwork = (int32_t)(n / p);
            if (p == 2) {
                s.kind = 1; s.mode = 2; s.mat_off = 0; s.sc_off = I.off(); s.flags = (h == 1) ? 1 : 0;
                for (int l = 0; l < L; ++l) {
                    const uint64_t q = pl->q[l], wpp = invmod(powmod(pl->root[l], pl->m / pp, q), q);
                    const uint64_t base = powmod(wpp, pp / (4 * h), q);
                    const uint64_t scale = (h == 1) ? invmod(mp % q, q) : 1;   // 2^-(e-1), applied once
                    for (int64_t r = 0; r < h; ++r) I.tab[l].push_back(sp(mulmod(powmod(base, 2 * r + 1, q), scale, q), q));
                    I.tab[l].push_back(sp(scale, q));                          // entry [h]: output-0 scale
                }
            } else {
                s.kind = 2; s.mode = 2; s.mat_off = I.off();
                for (int l = 0; l < L; ++l) {
                    const uint64_t q = pl->q[l], wpp = invmod(powmod(pl->root[l], pl->m / pp, q), q), wp = powmod(wpp, mp, q);
                    const uint64_t pinv = invmod(p % q, q);
                    for (int c = 0; c < p; ++c)
                        for (int r = 0; r < p; ++r) I.tab[l].push_back(sp(mulmod(powmod(wp, (uint64_t)c * r % p, q), pinv, q), q));
                }
                s.sc_off = I.off();
                for (int l = 0; l < L; ++l) {
                    const uint64_t q = pl->q[l], wpp = invmod(powmod(pl->root[l], pl->m / pp, q), q);
                    const uint64_t wph = powmod(wpp, pp / (p * h), q);
                    for (int64_t r = 0; r < h; ++r)
                        for (int c = 0; c < p; ++c) I.tab[l].push_back(sp(powmod(wph, (uint64_t)c * r % (p * h), q), q));
                }
            }
            I.steps.push_back(s);
            if (h == 1) break;
        }
        if (p > 2) {
            AbGenStep s{};
            s.kind = 0; s.P = p - 1; s.estride = (int32_t)rts; s.h = (int32_t)mp; s.sdiv = 1;
            s.mode = (e > 1) ? 1 : 0; s.nwork = (int32_t)(n / (p - 1));
            s.mat_off = I.off();
            for (int l = 0; l < L; ++l) {
                // a_j = (1/p) sum_{i=1}^{p-1} ahat_i (omega^(-i j) - omega^i)
                const uint64_t q = pl->q[l], wpp = powmod(pl->root[l], pl->m / pp, q), wp = powmod(wpp, mp, q);
                const uint64_t wpi = invmod(wp, q), pinv = invmod(p % q, q);
                for (int j0 = 0; j0 < p - 1; ++j0)
                    for (int i = 1; i <= p - 1; ++i) {
                        uint64_t t = (powmod(wpi, (uint64_t)i * j0 % p, q) + q - powmod(wp, i, q)) % q;
                        I.tab[l].push_back(sp(mulmod(t, pinv, q), q));
                    }
            }
            s.sc_off = I.off();
            if (e > 1)
                for (int l = 0; l < L; ++l) {
                    const uint64_t q = pl->q[l], wpp = invmod(powmod(pl->root[l], pl->m / pp, q), q);
                    for (int64_t j1 = 0; j1 < mp; ++j1)
                        for (int i0 = 0; i0 < p - 1; ++i0)
                            I.tab[l].push_back(sp(powmod(wpp, (uint64_t)digit_rev(p, e - 1, j1) * (i0 + 1) % pp, q), q));
                }
            I.steps.push_back(s);
        }
    }
    pl->steps_fwd = F.steps;
    pl->steps_inv = I.steps;
    pl->d_steps_fwd = dev_upload(pl, F.steps);
    pl->d_steps_inv = dev_upload(pl, I.steps);
    for (int l = 0; l < L; ++l) {
        pl->limb[l].gen_fwd = dev_upload(pl, F.tab[l]);
        pl->limb[l].gen_inv = dev_upload(pl, I.tab[l]);
    }
}

// inverse modulo any q >= 2 (extended Euclid), 0 when gcd(a, q) != 1
static uint64_t inv_or_zero(uint64_t a, uint64_t q) {
    int64_t r0 = (int64_t)q, r1 = (int64_t)(a % q), t0 = 0, t1 = 1;
    while (r1 != 0) {
        const int64_t k = r0 / r1;
        int64_t t = r0 - k * r1; r0 = r1; r1 = t;
        t = t0 - k * t1; t0 = t1; t1 = t;
    }
    if (r0 != 1) return 0;
    return (uint64_t)((t0 % (int64_t)q + (int64_t)q) % (int64_t)q);
}

// ------------------------------------------------------------------------------------ plan
static void plan_build(ab_plan *pl, const ab_prime_exp *pe, int32_t npe, const int64_t *moduli, int32_t L,
                       const int64_t *roots, bool want_crt) {
    AB_REQUIRE(pe && npe >= 0 && moduli && L >= 1 && L <= AB_MAX_LIMBS, AB_ERR_ARG, "bad plan arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        throw ab_error{AB_ERR_CUDA, "no CUDA device: this backend has no CPU fallback"};
    }
    AB_CHECK_CUDA(cudaGetDevice(&pl->device));
    cudaDeviceProp prop;
    AB_CHECK_CUDA(cudaGetDeviceProperties(&prop, pl->device));
    pl->sm_count = prop.multiProcessorCount;
    pl->axes = make_axes(pe, npe);
    pl->m = 1;
    pl->n = 1;
    for (auto &a : pl->axes) {
        pl->m *= a.pp;
        pl->n *= a.phi;
    }
    AB_REQUIRE(pl->n < (1ll << 24), AB_ERR_UNSUPPORTED, "ring dimension too large");
    pl->L = L;
    pl->pow2 = (pl->axes.size() == 1 && pl->axes[0].p == 2) || pl->axes.empty();
    pl->logn = 0;
    while ((1ll << pl->logn) < pl->n) ++pl->logn;
    pl->limb.assign(L, AbLimbDev{});
    pl->has_crt = want_crt;
    for (int l = 0; l < L; ++l) {
        const int64_t q = moduli[l];
        if (want_crt) {
            AB_REQUIRE(q > 2 && q < (1ll << 31) && is_prime((uint64_t)q), AB_ERR_ARG, "modulus must be an odd prime below 2^31");
            for (auto &a : pl->axes) AB_REQUIRE(q != a.p, AB_ERR_ARG, "modulus divides m");
        } else {
            // plaintext rings: any modulus (Z_7 in examples/Arithmetic.hs, Z_{2^k} in examples/HomomRLWR.hs)
            AB_REQUIRE(q >= 2 && q < (1ll << 31), AB_ERR_ARG, "modulus must lie in [2, 2^31)");
        }
        pl->q.push_back(q);
        AbLimbDev &ld = pl->limb[l];
        ld.q = (uint32_t)q;
        ld.mu = (uint64_t)(((unsigned __int128)1 << 64) / (uint64_t)q);
        for (size_t a = 0; a < pl->axes.size(); ++a) {
            uint64_t pi = inv_or_zero((uint64_t)pl->axes[a].p % (uint64_t)q, (uint64_t)q);   // 0: p not invertible mod q (divG refuses)
            ld.pinv[a] = (uint32_t)pi;
            ld.pinv_s[a] = shoup(pi, q);
        }
        if (want_crt) {
            if ((q - 1) % pl->m != 0) throw ab_error{AB_ERR_NO_ROOT, "m does not divide q-1 for modulus " + std::to_string(q)};
            uint64_t w = roots ? (uint64_t)roots[l] : powmod(smallest_generator(q), (q - 1) / pl->m, q);
            // must be a primitive m-th root of unity
            bool ok = powmod(w, pl->m, q) == 1;
            for (auto &a : pl->axes) ok = ok && powmod(w, pl->m / a.p, q) != 1;
            AB_REQUIRE(ok, AB_ERR_ARG, "supplied root is not a primitive m-th root of unity");
            pl->root.push_back((int64_t)w);
            int64_t mhat = (pl->m % 2 == 0) ? pl->m / 2 : pl->m;
            pl->mhat_inv.push_back((int64_t)invmod(mhat % q, q));
        } else {
            pl->root.push_back(0);
            pl->mhat_inv.push_back(0);
        }
    }
    if (!want_crt) return;
    if (pl->pow2 && pl->n >= 2)
        for (int l = 0; l < L; ++l) build_pow2_tables(pl, l, pl->limb[l]);
    build_generic(pl);
    ab_build_mixed(pl);
    // g at every CRT row
    std::vector<int64_t> units = crt_units(pl);
    for (int l = 0; l < L; ++l) {
        const uint64_t q = pl->q[l];
        std::vector<uint2_> g(pl->n), gi(pl->n);
        std::vector<std::vector<uint64_t>> one_minus;   // per odd prime: 1 - omega_p^k
        std::vector<int> primes;
        for (auto &a : pl->axes)
            if (a.p != 2) {
                uint64_t wp = powmod(pl->root[l], pl->m / a.p, q);
                std::vector<uint64_t> t(a.p);
                for (int k = 0; k < a.p; ++k) t[k] = (1 + q - powmod(wp, k, q)) % q;
                one_minus.push_back(t);
                primes.push_back(a.p);
            }
        for (int64_t i = 0; i < pl->n; ++i) {
            uint64_t v = 1;
            for (size_t k = 0; k < primes.size(); ++k) v = mulmod(v, one_minus[k][units[i] % primes[k]], q);
            g[i] = sp(v, q);
            gi[i] = sp(invmod(v, q), q);
        }
        pl->limb[l].g_crt = dev_upload(pl, g);
        pl->limb[l].ginv_crt = dev_upload(pl, gi);
    }
}

static int plan_create(ab_plan **out, const ab_prime_exp *pe, int32_t npe, const int64_t *moduli, int32_t L,
                       const int64_t *roots, bool want_crt) {
    ab_plan *pl = nullptr;
    int rc = ab_guard([&] {
        AB_REQUIRE(out, AB_ERR_ARG, "null out pointer");
        pl = new ab_plan();
        plan_build(pl, pe, npe, moduli, L, roots, want_crt);
        *out = pl;
    });
    if (rc != AB_OK && pl) {
        for (void *d : pl->owned) cudaFree(d);
        delete pl;
    }
    return rc;
}

extern "C" int ab_plan_create(ab_plan **out, const ab_prime_exp *pe, int32_t npe, const int64_t *moduli, int32_t L,
                              const int64_t *roots) {
    return plan_create(out, pe, npe, moduli, L, roots, true);
}
extern "C" int ab_plan_create_nocrt(ab_plan **out, const ab_prime_exp *pe, int32_t npe, const int64_t *moduli, int32_t L) {
    return plan_create(out, pe, npe, moduli, L, nullptr, false);
}
extern "C" void ab_plan_destroy(ab_plan *pl) {
    if (!pl) return;
    for (void *d : pl->owned) cudaFree(d);
    delete pl;
}
extern "C" int64_t ab_plan_m(const ab_plan *p) { return p ? p->m : 0; }
extern "C" int64_t ab_plan_n(const ab_plan *p) { return p ? p->n : 0; }
extern "C" int32_t ab_plan_limbs(const ab_plan *p) { return p ? p->L : 0; }
extern "C" int64_t ab_plan_root(const ab_plan *p, int32_t l) { return (p && l >= 0 && l < p->L) ? p->root[l] : 0; }

// ------------------------------------------------------------------------------------ extension tables
static int axis_of(const ab_plan *pl, int p) {
    for (size_t a = 0; a < pl->axes.size(); ++a)
        if (pl->axes[a].p == p) return (int)a;
    return -1;
}

extern "C" int ab_ext_create(ab_ext **out, const ab_plan *lo, const ab_plan *hi) {
    ab_ext *ex = nullptr;
    int rc = ab_guard([&] {
        AB_REQUIRE(out && lo && hi, AB_ERR_ARG, "null argument");
        AB_REQUIRE(hi->m % lo->m == 0, AB_ERR_ARG, "m does not divide m'");
        AB_REQUIRE(lo->L == hi->L && lo->q == hi->q, AB_ERR_ARG, "plans must carry the same moduli");
        ex = new ab_ext();
        ex->lo = lo; ex->hi = hi; ex->n_lo = lo->n; ex->n_hi = hi->n; ex->rel = hi->n / lo->n;
        const int A = (int)hi->axes.size();
        // per hi-axis: exponent in lo (0 if absent), totients
        std::vector<int> elo(A);
        std::vector<int64_t> dlo(A), dhi(A);
        for (int a = 0; a < A; ++a) {
            int b = axis_of(lo, hi->axes[a].p);
            elo[a] = b < 0 ? 0 : lo->axes[b].e;
            AB_REQUIRE(elo[a] <= hi->axes[a].e, AB_ERR_ARG, "m does not divide m'");
            dlo[a] = b < 0 ? 1 : lo->axes[b].phi;
            dhi[a] = hi->axes[a].phi;
        }
        const int64_t nlo = lo->n, nhi = hi->n, rel = ex->rel;
        std::vector<int32_t> ext_pow(nlo), pow_to_lo(nhi, -1), base_crt(nhi), above(nlo * rel), coeff(rel * nlo);
        std::vector<int32_t> edec_idx(nhi, -1);
        std::vector<int8_t> edec_sgn(nhi, 0);
        for (int64_t j = 0; j < nlo; ++j) {
            int64_t rem = j, idx = 0, mul = 1;
            for (int a = A - 1; a >= 0; --a) {
                int64_t ja = rem % dlo[a];
                rem /= dlo[a];
                idx += ja * mul;
                mul *= dhi[a];
            }
            ext_pow[j] = (int32_t)idx;
            pow_to_lo[idx] = (int32_t)j;
        }
        std::vector<int32_t> fill(nlo, 0);
        for (int64_t i = 0; i < nhi; ++i) {
            int64_t rem = i, idx = 0, mul = 1;
            // embedDec: per new odd prime the axis entry (j1 = 0, j0 in {0,1}) carries +x / -x
            int64_t didx = 0, dmul = 1;
            int sgn = 1;
            bool dz = false;
            for (int a = A - 1; a >= 0; --a) {
                const AbAxis &ax = hi->axes[a];
                int64_t ia = rem % dhi[a];
                rem /= dhi[a];
                int64_t i1 = ia / (ax.p - 1), i0 = ia % (ax.p - 1);
                if (elo[a] > 0) {
                    int64_t mplo = ipow(ax.p, elo[a] - 1);
                    idx += ((i1 % mplo) * (ax.p - 1) + i0) * mul;
                    mul *= dlo[a];
                    if (i1 < mplo) { didx += ia * dmul; } else dz = true;
                    dmul *= dlo[a];
                } else {
                    if (ia == 0) { /* +x */ } else if (ia == 1 && ax.p > 2) sgn = -sgn; else dz = true;
                }
            }
            base_crt[i] = (int32_t)idx;
            above[idx * rel + fill[idx]++] = (int32_t)i;
            if (!dz) { edec_idx[i] = (int32_t)didx; edec_sgn[i] = (int8_t)sgn; }
        }
        for (int64_t k = 0; k < rel; ++k)
            for (int64_t j = 0; j < nlo; ++j) {
                int64_t remj = j, remk = k, idx = 0, mul = 1;
                for (int a = A - 1; a >= 0; --a) {
                    const AbAxis &ax = hi->axes[a];
                    int64_t dk = dhi[a] / dlo[a];
                    int64_t ja = remj % dlo[a], ka = remk % dk;
                    remj /= dlo[a];
                    remk /= dk;
                    int64_t ia;
                    if (elo[a] == 0) ia = ka;
                    else ia = (ja / (ax.p - 1) + ipow(ax.p, elo[a] - 1) * ka) * (ax.p - 1) + ja % (ax.p - 1);
                    idx += ia * mul;
                    mul *= dhi[a];
                }
                coeff[k * nlo + j] = (int32_t)idx;
            }
        auto up32 = [&](const std::vector<int32_t> &v) {
            int32_t *d;
            AB_CHECK_CUDA(cudaMalloc((void **)&d, std::max<size_t>(1, v.size()) * 4));
            ex->owned.push_back(d);
            AB_CHECK_CUDA(cudaMemcpy(d, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
            return d;
        };
        ex->d_ext_pow = up32(ext_pow);
        ex->d_pow_to_lo = up32(pow_to_lo);
        ex->d_base_crt = up32(base_crt);
        ex->d_above_crt = up32(above);
        ex->d_coeff = up32(coeff);
        ex->d_edec_idx = up32(edec_idx);
        {
            int8_t *d;
            AB_CHECK_CUDA(cudaMalloc((void **)&d, std::max<size_t>(1, edec_sgn.size())));
            ex->owned.push_back(d);
            AB_CHECK_CUDA(cudaMemcpy(d, edec_sgn.data(), edec_sgn.size(), cudaMemcpyHostToDevice));
            ex->d_edec_sgn = d;
        }
        // twaceCRT weights (only if both plans have CRT tables)
        if (lo->has_crt && hi->has_crt) {
            std::vector<uint2_> w((size_t)lo->L * nhi);
            std::vector<uint2_> ghi(nhi), gilo(nlo);
            for (int l = 0; l < lo->L; ++l) {
                const uint64_t q = lo->q[l];
                AB_CHECK_CUDA(cudaMemcpy(ghi.data(), hi->limb[l].g_crt, nhi * sizeof(uint2_), cudaMemcpyDeviceToHost));
                AB_CHECK_CUDA(cudaMemcpy(gilo.data(), lo->limb[l].ginv_crt, nlo * sizeof(uint2_), cudaMemcpyDeviceToHost));
                int64_t mh_lo = (lo->m % 2 == 0) ? lo->m / 2 : lo->m;
                uint64_t scale = mulmod(mh_lo % q, (uint64_t)hi->mhat_inv[l], q);
                for (int64_t i = 0; i < nhi; ++i)
                    w[(size_t)l * nhi + i] = sp(mulmod(mulmod(ghi[i].x, gilo[base_crt[i]].x, q), scale, q), q);
            }
            uint2_ *d;
            AB_CHECK_CUDA(cudaMalloc((void **)&d, w.size() * sizeof(uint2_)));
            ex->owned.push_back(d);
            AB_CHECK_CUDA(cudaMemcpy(d, w.data(), w.size() * sizeof(uint2_), cudaMemcpyHostToDevice));
            ex->d_tw_weight = d;
        }
        *out = ex;
    });
    if (rc != AB_OK && ex) {
        for (void *d : ex->owned) cudaFree(d);
        delete ex;
    }
    return rc;
}
extern "C" void ab_ext_destroy(ab_ext *ex) {
    if (!ex) return;
    for (void *d : ex->owned) cudaFree(d);
    delete ex;
}
extern "C" int64_t ab_ext_rel_dim(const ab_ext *e) { return e ? e->rel : 0; }
