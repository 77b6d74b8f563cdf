// ab_abi.cpp -- extern "C" entry points for the Tensor methods (include/alchemy_b200.h).
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>

#include "ab_internal.h"
#include "ab_kernels.h"

static std::atomic<uint64_t> g_launches{0};
void ab_count_launch(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
extern "C" uint64_t ab_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }
void ab_check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) throw ab_error{AB_ERR_CUDA, std::string(what) + " launch: " + cudaGetErrorString(e)};
}

static AbView make_view(const ab_plan *p, int64_t *y, int64_t batch, int32_t limb0, int32_t nlimbs, int layout) {
    AB_REQUIRE(p != nullptr, AB_ERR_ARG, "null plan");
    AB_REQUIRE(y != nullptr || batch == 0, AB_ERR_ARG, "null buffer");
    AB_REQUIRE(batch >= 0, AB_ERR_ARG, "negative batch");
    AB_REQUIRE(limb0 >= 0 && nlimbs >= 1 && limb0 + nlimbs <= p->L, AB_ERR_ARG, "limb range outside the plan");
    AB_REQUIRE(layout == AB_LAYOUT_PLANAR || layout == AB_LAYOUT_INTERLEAVED, AB_ERR_ARG, "unknown layout");
    return AbView{y, batch, limb0, nlimbs, layout};
}

void ab_crt_dispatch(const ab_plan *p, const AbView &v, bool inverse, cudaStream_t st) {
    AB_REQUIRE(p->has_crt, AB_ERR_NO_ROOT, "plan was created without CRT tables");
    if (p->n == 1) return;   // m in {1, 2}: CRT is the identity
    if (ab_pow2_supported(p) && (v.layout == AB_LAYOUT_PLANAR || v.nlimbs == 1)) ab_launch_crt_pow2(p, v, inverse, st);
    else if (ab_mixed_supported(p, v)) ab_launch_crt_mixed(p, v, inverse, st);
    else ab_launch_crt_generic(p, v, inverse, st);
}
void ab_crt_rows(const ab_plan *p, int64_t *y, int64_t rows, int32_t limb, bool inverse, cudaStream_t st) {
    ab_crt_dispatch(p, AbView{y, rows, limb, 1, AB_LAYOUT_PLANAR}, inverse, st);
}

extern "C" int ab_crt_dev(const ab_plan *p, int64_t *y, int64_t batch, int32_t limb0, int32_t nlimbs, int layout, void *stream) {
    return ab_guard([&] { ab_crt_dispatch(p, make_view(p, y, batch, limb0, nlimbs, layout), false, (cudaStream_t)stream); });
}
extern "C" int ab_crt_inv_dev(const ab_plan *p, int64_t *y, int64_t batch, int32_t limb0, int32_t nlimbs, int layout, void *stream) {
    return ab_guard([&] { ab_crt_dispatch(p, make_view(p, y, batch, limb0, nlimbs, layout), true, (cudaStream_t)stream); });
}

#define AB_PRIME_ENTRY(name, op)                                                                                              \
    extern "C" int name(const ab_plan *p, int64_t *y, int64_t batch, int32_t limb0, int32_t nlimbs, int layout, void *stream) { \
        return ab_guard([&] { ab_launch_prime_op(p, op, make_view(p, y, batch, limb0, nlimbs, layout), (cudaStream_t)stream); }); \
    }
AB_PRIME_ENTRY(ab_l_dev, AB_PR_L)
AB_PRIME_ENTRY(ab_l_inv_dev, AB_PR_LINV)
AB_PRIME_ENTRY(ab_mul_g_pow_dev, AB_PR_GPOW)
AB_PRIME_ENTRY(ab_mul_g_dec_dev, AB_PR_GDEC)
AB_PRIME_ENTRY(ab_div_g_pow_dev, AB_PR_GINVPOW)
AB_PRIME_ENTRY(ab_div_g_dec_dev, AB_PR_GINVDEC)

extern "C" int ab_mul_g_crt_dev(const ab_plan *p, int64_t *y, int64_t batch, int32_t limb0, int32_t nlimbs, int layout, void *stream) {
    return ab_guard([&] { ab_launch_elementwise(p, AB_EW_GCRT, make_view(p, y, batch, limb0, nlimbs, layout), nullptr, 0, nullptr, (cudaStream_t)stream); });
}
extern "C" int ab_div_g_crt_dev(const ab_plan *p, int64_t *y, int64_t batch, int32_t limb0, int32_t nlimbs, int layout, void *stream) {
    return ab_guard([&] { ab_launch_elementwise(p, AB_EW_GINVCRT, make_view(p, y, batch, limb0, nlimbs, layout), nullptr, 0, nullptr, (cudaStream_t)stream); });
}
extern "C" int ab_mul_dev(const ab_plan *p, int64_t *y, const int64_t *x, int64_t batch, int64_t x_batch, int32_t limb0, int32_t nlimbs, int layout, void *stream) {
    return ab_guard([&] { ab_launch_elementwise(p, AB_EW_MUL, make_view(p, y, batch, limb0, nlimbs, layout), x, x_batch, nullptr, (cudaStream_t)stream); });
}
extern "C" int ab_add_dev(const ab_plan *p, int64_t *y, const int64_t *x, int64_t batch, int64_t x_batch, int32_t limb0, int32_t nlimbs, int layout, void *stream) {
    return ab_guard([&] { ab_launch_elementwise(p, AB_EW_ADD, make_view(p, y, batch, limb0, nlimbs, layout), x, x_batch, nullptr, (cudaStream_t)stream); });
}
extern "C" int ab_sub_dev(const ab_plan *p, int64_t *y, const int64_t *x, int64_t batch, int64_t x_batch, int32_t limb0, int32_t nlimbs, int layout, void *stream) {
    return ab_guard([&] { ab_launch_elementwise(p, AB_EW_SUB, make_view(p, y, batch, limb0, nlimbs, layout), x, x_batch, nullptr, (cudaStream_t)stream); });
}
extern "C" int ab_neg_dev(const ab_plan *p, int64_t *y, int64_t batch, int32_t limb0, int32_t nlimbs, int layout, void *stream) {
    return ab_guard([&] { ab_launch_elementwise(p, AB_EW_NEG, make_view(p, y, batch, limb0, nlimbs, layout), nullptr, 0, nullptr, (cudaStream_t)stream); });
}
extern "C" int ab_scalar_mul_dev(const ab_plan *p, int64_t *y, const int64_t *scalars, int64_t batch, int32_t limb0, int32_t nlimbs, int layout, void *stream) {
    return ab_guard([&] { ab_launch_elementwise(p, AB_EW_SCALAR, make_view(p, y, batch, limb0, nlimbs, layout), nullptr, 0, scalars, (cudaStream_t)stream); });
}

// ------------------------------------------------------------------------------------ extension rings
static void ext_args(const ab_ext *e, const void *out, const void *in, int64_t batch, int32_t nlimbs) {
    AB_REQUIRE(e != nullptr, AB_ERR_ARG, "null ext");
    AB_REQUIRE((out && in) || batch == 0, AB_ERR_ARG, "null buffer");
    AB_REQUIRE(batch >= 0 && nlimbs >= 1 && nlimbs <= e->lo->L, AB_ERR_ARG, "bad batch or limb count");
}
extern "C" int ab_twace_pow_dec_dev(const ab_ext *e, int64_t *out_lo, const int64_t *in_hi, int64_t batch, int32_t nlimbs, void *stream) {
    return ab_guard([&] {
        ext_args(e, out_lo, in_hi, batch, nlimbs);
        ab_launch_index_map(0, out_lo, in_hi, e->d_ext_pow, nullptr, e->lo, batch * nlimbs, e->n_lo, e->n_hi, 1, nlimbs, (cudaStream_t)stream);
    });
}
extern "C" int ab_embed_pow_dev(const ab_ext *e, int64_t *out_hi, const int64_t *in_lo, int64_t batch, int32_t nlimbs, void *stream) {
    return ab_guard([&] {
        ext_args(e, out_hi, in_lo, batch, nlimbs);
        ab_launch_index_map(0, out_hi, in_lo, e->d_pow_to_lo, nullptr, e->lo, batch * nlimbs, e->n_hi, e->n_lo, 1, nlimbs, (cudaStream_t)stream);
    });
}
extern "C" int ab_embed_dec_dev(const ab_ext *e, int64_t *out_hi, const int64_t *in_lo, int64_t batch, int32_t nlimbs, void *stream) {
    return ab_guard([&] {
        ext_args(e, out_hi, in_lo, batch, nlimbs);
        ab_launch_index_map(0, out_hi, in_lo, e->d_edec_idx, e->d_edec_sgn, e->lo, batch * nlimbs, e->n_hi, e->n_lo, 1, nlimbs, (cudaStream_t)stream);
    });
}
extern "C" int ab_embed_crt_dev(const ab_ext *e, int64_t *out_hi, const int64_t *in_lo, int64_t batch, int32_t nlimbs, void *stream) {
    return ab_guard([&] {
        ext_args(e, out_hi, in_lo, batch, nlimbs);
        ab_launch_index_map(0, out_hi, in_lo, e->d_base_crt, nullptr, e->lo, batch * nlimbs, e->n_hi, e->n_lo, 1, nlimbs, (cudaStream_t)stream);
    });
}
extern "C" int ab_twace_crt_dev(const ab_ext *e, int64_t *out_lo, const int64_t *in_hi, int64_t batch, int32_t nlimbs, void *stream) {
    return ab_guard([&] {
        ext_args(e, out_lo, in_hi, batch, nlimbs);
        AB_REQUIRE(e->d_tw_weight != nullptr, AB_ERR_NO_ROOT, "twaceCRT needs CRT tables on both plans");
        ab_launch_twace_crt(e, out_lo, in_hi, batch, nlimbs, (cudaStream_t)stream);
    });
}
extern "C" int ab_coeffs_dev(const ab_ext *e, int64_t *out_lo, const int64_t *in_hi, int64_t batch, int32_t nlimbs, void *stream) {
    return ab_guard([&] {
        ext_args(e, out_lo, in_hi, batch, nlimbs);
        ab_launch_index_map(1, out_lo, in_hi, e->d_coeff, nullptr, e->lo, batch * e->rel * nlimbs, e->n_lo, e->n_hi, e->rel, nlimbs, (cudaStream_t)stream);
    });
}
extern "C" int ab_uncoeffs_dev(const ab_ext *e, int64_t *out_hi, const int64_t *in_lo, int64_t batch, int32_t nlimbs, void *stream) {
    return ab_guard([&] {
        ext_args(e, out_hi, in_lo, batch, nlimbs);
        ab_launch_index_map(2, out_hi, in_lo, e->d_coeff, nullptr, e->lo, batch * e->rel * nlimbs, e->n_hi, e->n_lo, e->rel, nlimbs, (cudaStream_t)stream);
    });
}

// ------------------------------------------------------------------------------------ cross-limb
extern "C" int32_t ab_gadget_len(const ab_plan *p, int32_t nlimbs, int32_t base) {
    if (!p || nlimbs < 1 || nlimbs > p->L || base < 0 || base == 1) return -1;
    int32_t nd = 0;
    for (int l = 0; l < nlimbs; ++l) nd += ab_gadget_digits(p->q[l], base);
    return nd;
}
extern "C" int ab_decompose_dev(const ab_plan *p, int64_t *out, const int64_t *in, int64_t batch, int32_t nlimbs, int32_t out_limbs, int32_t base, void *stream) {
    return ab_guard([&] {
        AB_REQUIRE(p && ((out && in) || batch == 0), AB_ERR_ARG, "null argument");
        AB_REQUIRE(batch >= 0 && nlimbs >= 1 && nlimbs <= p->L && out_limbs >= 1 && out_limbs <= p->L, AB_ERR_ARG, "bad limb counts");
        AB_REQUIRE(base == 0 || base >= 2, AB_ERR_ARG, "gadget base must be 0 (TrivGad) or >= 2");
        ab_launch_decompose(p, out, in, batch, nlimbs, out_limbs, base, (cudaStream_t)stream);
    });
}
extern "C" int ab_rescale_drop_dev(const ab_plan *p, int64_t *out, const int64_t *in, int64_t batch, int32_t nlimbs, int32_t drop, void *stream) {
    return ab_guard([&] {
        AB_REQUIRE(p && ((out && in) || batch == 0), AB_ERR_ARG, "null argument");
        AB_REQUIRE(batch >= 0 && nlimbs >= 2 && nlimbs <= p->L && drop >= 0 && drop < nlimbs, AB_ERR_ARG, "bad limb arguments");
        AB_REQUIRE(out != in || drop == nlimbs - 1, AB_ERR_ARG, "in-place rescale only when dropping the last limb");
        ab_launch_rescale_drop(p, out, in, batch, nlimbs, drop, (cudaStream_t)stream);
    });
}

// ------------------------------------------------------------------------------------ host-buffer entries
// What a `foreign import ccall` from Lol's CPP.hs binds today: the buffer lives in Haskell-managed
// host memory.  Copy in, run on the device, copy out, synchronise.
static void run_op(const ab_plan *p, int op, const AbView &v, cudaStream_t st) {
    switch (op) {
        case AB_OP_CRT: ab_crt_dispatch(p, v, false, st); break;
        case AB_OP_CRT_INV: ab_crt_dispatch(p, v, true, st); break;
        case AB_OP_L: ab_launch_prime_op(p, AB_PR_L, v, st); break;
        case AB_OP_L_INV: ab_launch_prime_op(p, AB_PR_LINV, v, st); break;
        case AB_OP_MUL_G_POW: ab_launch_prime_op(p, AB_PR_GPOW, v, st); break;
        case AB_OP_MUL_G_DEC: ab_launch_prime_op(p, AB_PR_GDEC, v, st); break;
        case AB_OP_DIV_G_POW: ab_launch_prime_op(p, AB_PR_GINVPOW, v, st); break;
        case AB_OP_DIV_G_DEC: ab_launch_prime_op(p, AB_PR_GINVDEC, v, st); break;
        case AB_OP_MUL_G_CRT: ab_launch_elementwise(p, AB_EW_GCRT, v, nullptr, 0, nullptr, st); break;
        case AB_OP_DIV_G_CRT: ab_launch_elementwise(p, AB_EW_GINVCRT, v, nullptr, 0, nullptr, st); break;
        default: throw ab_error{AB_ERR_ARG, "unknown op"};
    }
}

// stream-ordered scratch from the device's default memory pool (kept warm: the release threshold is raised
// once, so repeated host-entry calls do not pay cudaMalloc/cudaFree each time)
struct DevBuf {
    int64_t *p = nullptr;
    cudaStream_t st;
    DevBuf(size_t bytes, cudaStream_t s) : st(s) {
        static std::once_flag once;
        std::call_once(once, [] {
            int dev = 0;
            cudaMemPool_t pool;
            if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                uint64_t keep = UINT64_MAX;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            cudaGetLastError();
        });
        AB_CHECK_CUDA(cudaMallocAsync((void **)&p, bytes ? bytes : 8, st));
    }
    ~DevBuf() { cudaFreeAsync(p, st); }
    DevBuf(const DevBuf &) = delete;
};

// Host-buffer calls are pipelined: the batch is cut into chunks of about 16 MB (AB_HOST_CHUNK_MB; measured at n = 32768, L = 4 with
// four streams: 2 MB 114 k, 4 MB 126 k, 8 MB 150 k, 16 MB 162 k, 32 MB 145 k transforms/s end to end, i.e. 42.5 GB/s per direction of
// the 49.6 GB/s the box reaches with both directions busy, tools/micro/pcie.py) that rotate over four
// internal streams (H2D copy, kernels, D2H copy of one chunk are stream-ordered; chunks on different streams
// overlap), so the two PCIe directions and the kernels run concurrently.  With pinned host memory the copies
// are truly asynchronous; with pageable memory the runtime stages them and the call still returns the same data.
struct HostPipe {
    static constexpr int kStreams = 4;
    cudaStream_t st[kStreams] = {nullptr, nullptr, nullptr, nullptr};
    int device = -1;
    void ensure() {
        int dev = 0;
        AB_CHECK_CUDA(cudaGetDevice(&dev));
        if (device == dev) return;
        for (auto &s : st) {
            if (s) cudaStreamDestroy(s);
            AB_CHECK_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        }
        device = dev;
    }
};
static thread_local HostPipe g_pipe;

extern "C" int ab_tensor_op_host(const ab_plan *p, int op, int64_t *y_host, int64_t batch, int layout) {
    return ab_guard([&] {
        AB_REQUIRE(p && (y_host || batch == 0) && batch >= 0, AB_ERR_ARG, "bad arguments");
        AB_REQUIRE(layout == AB_LAYOUT_PLANAR || layout == AB_LAYOUT_INTERLEAVED, AB_ERR_ARG, "unknown layout");
        if (batch == 0) return;
        g_pipe.ensure();
        const size_t item = (size_t)p->L * p->n * sizeof(int64_t);
        static const size_t chunk_mb = getenv("AB_HOST_CHUNK_MB") && atoi(getenv("AB_HOST_CHUNK_MB")) > 0 ? (size_t)atoi(getenv("AB_HOST_CHUNK_MB")) : 16;      // experiments
        const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(batch, (int64_t)((chunk_mb << 20) / item)));
        const bool relayout = layout == AB_LAYOUT_INTERLEAVED && p->L > 1;
        const int ns = (int)std::min<int64_t>(HostPipe::kStreams, (batch + chunk - 1) / chunk);
        std::vector<std::unique_ptr<DevBuf>> d(ns), t(ns);
        for (int i = 0; i < ns; ++i) {
            d[i].reset(new DevBuf(chunk * item, g_pipe.st[i]));
            if (relayout) t[i].reset(new DevBuf(chunk * item, g_pipe.st[i]));
        }
        int64_t k = 0;
        for (int64_t b0 = 0; b0 < batch; b0 += chunk, ++k) {
            const int i = (int)(k % ns);
            cudaStream_t st = g_pipe.st[i];
            const int64_t nb = std::min(chunk, batch - b0);
            int64_t *h = y_host + b0 * (int64_t)(item / sizeof(int64_t));
            AB_CHECK_CUDA(cudaMemcpyAsync(d[i]->p, h, nb * item, cudaMemcpyHostToDevice, st));
            if (relayout) {
                // de-interleave once on the device so every kernel streams contiguous limbs
                ab_launch_layout(t[i]->p, d[i]->p, nb, p->L, p->n, true, st);
                run_op(p, op, AbView{t[i]->p, nb, 0, p->L, AB_LAYOUT_PLANAR}, st);
                ab_launch_layout(d[i]->p, t[i]->p, nb, p->L, p->n, false, st);
            } else {
                run_op(p, op, AbView{d[i]->p, nb, 0, p->L, AB_LAYOUT_PLANAR}, st);
            }
            AB_CHECK_CUDA(cudaMemcpyAsync(h, d[i]->p, nb * item, cudaMemcpyDeviceToHost, st));
        }
        d.clear();      // stream-ordered frees
        t.clear();
        for (int i = 0; i < ns; ++i) AB_CHECK_CUDA(cudaStreamSynchronize(g_pipe.st[i]));
    });
}
extern "C" int ab_mul_host(const ab_plan *p, int64_t *y_host, const int64_t *x_host, int64_t batch, int layout) {
    return ab_guard([&] {
        AB_REQUIRE(p && ((y_host && x_host) || batch == 0) && batch >= 0, AB_ERR_ARG, "bad arguments");
        if (batch == 0) return;
        const size_t bytes = (size_t)batch * p->L * p->n * sizeof(int64_t);
        cudaStream_t st = 0;
        DevBuf dy(bytes, st), dx(bytes, st);
        AB_CHECK_CUDA(cudaMemcpyAsync(dy.p, y_host, bytes, cudaMemcpyHostToDevice, st));
        AB_CHECK_CUDA(cudaMemcpyAsync(dx.p, x_host, bytes, cudaMemcpyHostToDevice, st));
        ab_launch_elementwise(p, AB_EW_MUL, AbView{dy.p, batch, 0, p->L, layout}, dx.p, batch, nullptr, st);
        AB_CHECK_CUDA(cudaMemcpyAsync(y_host, dy.p, bytes, cudaMemcpyDeviceToHost, st));
        AB_CHECK_CUDA(cudaStreamSynchronize(st));
    });
}

// ---- drop-in names with the upstream argument order [RECALL] --------------------------------
static std::mutex g_cache_mu;
static std::map<std::vector<int64_t>, ab_plan *> g_cache;

static const ab_plan *cached_plan(const ab_prime_exp *pe, int16_t npe, const int64_t *qs, int16_t L, const int64_t *const *ru, bool crt) {
    AB_REQUIRE(pe && qs && npe >= 0 && L >= 1, AB_ERR_ARG, "bad arguments");
    std::vector<int64_t> key;
    key.push_back(crt ? 1 : 0);
    for (int i = 0; i < npe; ++i) { key.push_back(pe[i].prime); key.push_back(pe[i].exponent); }
    key.push_back(-1);
    for (int l = 0; l < L; ++l) key.push_back(qs[l]);
    // omega_m per limb from the caller's per-prime-power root tables: ru[i] is the interleaved table of
    // powers of omega_{pp_i}, entry 1 being the root itself; omega_m = prod_i omega_{pp_i}^(c_i) with
    // c_i = (m/pp_i)^-1 mod pp_i (the identification zeta_m = prod zeta_pp^(c_i) of ab_plan.cpp).
    std::vector<int64_t> roots;
    if (crt && ru) {
        int64_t m = 1;
        for (int i = 0; i < npe; ++i) m *= abh::ipow(pe[i].prime, pe[i].exponent);
        for (int l = 0; l < L; ++l) {
            uint64_t w = 1;
            for (int i = 0; i < npe; ++i) {
                const int64_t pp = abh::ipow(pe[i].prime, pe[i].exponent), co = (m / pp) % pp;
                int64_t c = 0;
                for (int64_t x = 1; x < pp; ++x) if (co * x % pp == 1) { c = x; break; }
                if (pp == 2) c = 1;
                w = abh::mulmod(w, abh::powmod((uint64_t)ru[i][1 * L + l], (uint64_t)c, (uint64_t)qs[l]), (uint64_t)qs[l]);
            }
            roots.push_back((int64_t)w);
            key.push_back((int64_t)w);
        }
    }
    std::lock_guard<std::mutex> lk(g_cache_mu);
    auto it = g_cache.find(key);
    if (it != g_cache.end()) return it->second;
    ab_plan *pl = nullptr;
    int rc = crt ? ab_plan_create(&pl, pe, npe, qs, L, roots.empty() ? nullptr : roots.data()) : ab_plan_create_nocrt(&pl, pe, npe, qs, L);
    if (rc != AB_OK) throw ab_error{rc, ab_last_error()};
    g_cache[key] = pl;
    return pl;
}
static int compat(int op, int16_t tupSize, int64_t *y, const ab_prime_exp *pe, int16_t npe, const int64_t *qs, const int64_t *const *ru) {
    const ab_plan *pl = nullptr;
    const bool crt = (op == AB_OP_CRT || op == AB_OP_CRT_INV);
    int rc = ab_guard([&] { pl = cached_plan(pe, npe, qs, tupSize, ru, crt); });
    if (rc != AB_OK) return rc;
    return ab_tensor_op_host(pl, op, y, 1, AB_LAYOUT_INTERLEAVED);
}
extern "C" int tensorCRTRq(int16_t tupSize, int64_t *y, int64_t, const ab_prime_exp *pe, int16_t npe, const int64_t *const *ru, const int64_t *qs) {
    return compat(AB_OP_CRT, tupSize, y, pe, npe, qs, ru);
}
extern "C" int tensorCRTInvRq(int16_t tupSize, int64_t *y, int64_t, const ab_prime_exp *pe, int16_t npe, const int64_t *const *ruinv, const int64_t *, const int64_t *qs) {
    // upstream passes the tables of omega^-1; recover omega per prime power by inverting entry 1
    const ab_plan *pl = nullptr;
    int rc = ab_guard([&] {
        AB_REQUIRE(ruinv && pe && qs, AB_ERR_ARG, "bad arguments");
        std::vector<std::vector<int64_t>> tabs(npe);
        std::vector<const int64_t *> ptrs(npe);
        for (int i = 0; i < npe; ++i) {
            tabs[i].assign((size_t)2 * tupSize, 1);
            for (int l = 0; l < tupSize; ++l) tabs[i][tupSize + l] = (int64_t)abh::invmod((uint64_t)ruinv[i][tupSize + l], (uint64_t)qs[l]);
            ptrs[i] = tabs[i].data();
        }
        pl = cached_plan(pe, npe, qs, tupSize, ptrs.data(), true);
    });
    if (rc != AB_OK) return rc;
    return ab_tensor_op_host(pl, AB_OP_CRT_INV, y, 1, AB_LAYOUT_INTERLEAVED);
}
#define AB_COMPAT(name, op)                                                                                              \
    extern "C" int name(int16_t tupSize, int64_t *y, int64_t, const ab_prime_exp *pe, int16_t npe, const int64_t *qs) { \
        return compat(op, tupSize, y, pe, npe, qs, nullptr);                                                           \
    }
AB_COMPAT(tensorLRq, AB_OP_L)
AB_COMPAT(tensorLInvRq, AB_OP_L_INV)
AB_COMPAT(tensorGPowRq, AB_OP_MUL_G_POW)
AB_COMPAT(tensorGDecRq, AB_OP_MUL_G_DEC)
AB_COMPAT(tensorGInvPowRq, AB_OP_DIV_G_POW)
AB_COMPAT(tensorGInvDecRq, AB_OP_DIV_G_DEC)
extern "C" int mulRq(int16_t tupSize, int64_t *a, const int64_t *b, int64_t totm, const int64_t *qs) {
    // pointwise product needs no ring structure: a 1-axis dummy plan with n = totm is enough; use a
    // no-CRT plan over the prime power 2^1 and launch the streaming kernel directly on totm entries.
    return ab_guard([&] {
        AB_REQUIRE(a && b && qs && tupSize >= 1 && tupSize <= AB_MAX_LIMBS && totm >= 0, AB_ERR_ARG, "bad arguments");
        ab_prime_exp pe{2, 1};
        const ab_plan *pl = cached_plan(&pe, 1, qs, tupSize, nullptr, false);
        // n = 1 for that plan, so `batch` carries the coefficient count; interleaved tuples of size tupSize
        int rc = ab_mul_host(pl, a, b, totm, AB_LAYOUT_INTERLEAVED);
        if (rc != AB_OK) throw ab_error{rc, ab_last_error()};
    });
}
