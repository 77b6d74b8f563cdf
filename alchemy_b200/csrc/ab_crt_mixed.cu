// ab_crt_mixed.cu -- batched CRT / crtInv for arbitrary cyclotomic index m, prime moduli below 2^31 (sm_100a).
//
// Replaces lol-cpp `tensorCRTRq` / `tensorCRTInvRq` [RECALL] on the non-power-of-two rings of the reference's
// examples (examples/Common.hs:41-54: m' = 11648 ... 54600) and of BASELINE config 3 (m = 40320).
// One CTA owns one (ring element, limb): the n = phi(m) int64 residues are read once from HBM with 16-byte
// loads and narrowed to uint32 in shared memory, the program of ab_mixed_plan.cpp runs out of shared memory,
// and the result is written once: 16 n bytes of HBM traffic per transform, the algorithmic minimum.
//
// A dense step computes out_r = sum_c M[r][c] x_c for P <= 12 strided residues held in registers.  The
// constants are stored as M * 2^32 mod q, the sum is accumulated exactly in 64 bits (IMAD.WIDE, one
// instruction per term) and reduced ONCE per output with a Montgomery step: P + 4 instructions per output
// instead of 7 per term.  The limb's constant tables (a few KB) sit in shared memory and are fetched with
// 16-byte broadcast loads.  Radix-2 stages of the 2^e axis run three at a time on 8 residues per thread.
#define AB_MX_KERNEL_INV 0
#include "ab_crt_mixed_kernel.cuh"

void ab_mx_launch_inv(dim3 grid, int T, size_t smem, cudaStream_t st, const MxArgs &args, int64_t *y);   // ab_crt_mixed_inv.cu

bool ab_mixed_supported(const ab_plan *p, const AbView &v) {
    if (p->mx_ok.empty() || p->mx_fwd.empty() || p->mx_tab_fwd.empty()) return false;
    for (int l = 0; l < v.nlimbs; ++l)
        if (!p->mx_ok[v.limb0 + l]) return false;
    if (v.layout == AB_LAYOUT_PLANAR && (((uintptr_t)v.y & 15) != 0 || (p->n & 1))) return false;
    return true;
}

void ab_launch_crt_mixed(const ab_plan *p, const AbView &v, bool inverse, cudaStream_t st) {
    ab_launch_crt_mixed_fused(p, v, inverse, nullptr, st);
}

// `v` describes the DESTINATION (planar): v.batch = number of transforms per limb (units * nd for digit loads)
void ab_launch_crt_mixed_fused(const ab_plan *p, const AbView &v, bool inverse, const AbMxFused *fx, cudaStream_t st) {
    AB_REQUIRE(ab_mixed_supported(p, v), AB_ERR_UNSUPPORTED, "mixed kernel: plan not supported");
    AB_REQUIRE(fx == nullptr || v.layout == AB_LAYOUT_PLANAR, AB_ERR_UNSUPPORTED, "fused modes need the planar layout");
    if (v.batch == 0) return;
    MxArgs args{};
    if (fx) args.fx = *fx;
    if (args.fx.pre_l || args.fx.post_linv) {
        args.fx.n_odd = 0;
        for (const AbAxis &ax : p->axes)
            if (ax.p != 2) { args.fx.odd[args.fx.n_odd].d = ax.p - 1; args.fx.odd[args.fx.n_odd].rts = (int32_t)ax.rts; ++args.fx.n_odd; }
    }
    const auto &steps = inverse ? p->mx_inv : p->mx_fwd;
    args.nsteps = (int32_t)steps.size();
    for (size_t i = 0; i < steps.size(); ++i) args.steps[i] = steps[i];
    args.n = (int32_t)p->n;
    args.tab_words[0] = inverse ? p->mx_words_inv[0] : p->mx_words_fwd[0];
    args.tab_words[1] = inverse ? p->mx_words_inv[1] : p->mx_words_fwd[1];
    int max_words = 0;
    args.nlimbs = v.nlimbs;
    args.batch = v.batch;
    if (v.layout == AB_LAYOUT_PLANAR) { args.estride = 1; args.lstride = p->n; args.bstride = p->n * v.nlimbs; }
    else { args.estride = v.nlimbs; args.lstride = 1; args.bstride = p->n * v.nlimbs; }
    if (args.fx.dst_ustride) args.bstride = args.fx.dst_ustride;
    for (int l = 0; l < v.nlimbs; ++l) {
        const AbLimbDev &ld = p->limb[v.limb0 + l];
        args.limb[l].q = ld.q;
        args.limb[l].qinv_neg = neg_inv32(ld.q);
        args.limb[l].tab = inverse ? p->mx_tab_inv[v.limb0 + l] : p->mx_tab_fwd[v.limb0 + l];
        args.limb[l].wide = p->mx_ok[v.limb0 + l] == 2;
        max_words = std::max(max_words, args.tab_words[args.limb[l].wide]);
    }
    const size_t smem = ((size_t)max_words + (size_t)p->n) * sizeof(uint32_t);
    AB_REQUIRE(smem <= 227 * 1024, AB_ERR_UNSUPPORTED, "ring too large for the shared-memory transform");
    // one-dimensional grid over (ring element, limb); grid.x is limited to 2^31-1
    AB_REQUIRE(v.batch * v.nlimbs < (1ll << 31), AB_ERR_ARG, "batch too large for one launch");
    args.interleave = ab_mx_interleave();
    const dim3 grid((unsigned)(v.batch * v.nlimbs));
    static const int force_t = getenv("AB_MX_T") ? atoi(getenv("AB_MX_T")) : 0;      // experiments only
    const int T = force_t ? force_t : pick_threads(steps, (int)p->n, smem, 256 * AB_MX_MINB);
    if (inverse) ab_mx_launch_inv(grid, T, smem, st, args, v.y);
    else ab_mx_launch_fwd(grid, T, smem, st, args, v.y);
    ab_count_launch();
    ab_check_launch("crt_mixed");
}

