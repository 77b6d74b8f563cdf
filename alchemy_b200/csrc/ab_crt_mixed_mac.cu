// ab_crt_mixed_mac.cu -- the fused gadget inner product of keySwitchQuadCirc (Eval.hs:133) and tunnel (Eval.hs:134):
// digit extraction, forward CRT and multiply-accumulate against the prepared hint in one kernel (sm_100a).
#define AB_MX_MAC_SMEM 0
#include "ab_crt_mixed_mac_kernel.cuh"

void ab_mx_launch_mac_smem(dim3 grid, int T, size_t smem, cudaStream_t st, const MxArgs &args, const AbMxMac &mac);   // ab_crt_mixed_mac_smem.cu

// out_c = add_c + sum over the parts of a unit of the partial sums (parts > 1 only)
__global__ void __launch_bounds__(256) k_mac_combine(const __grid_constant__ MxArgs args, const __grid_constant__ AbMxMac mac) {
    const int n = mac.n_out > 0 ? mac.n_out : args.n, L = args.nlimbs, parts = mac.parts;
    const int64_t per = (int64_t)L * n, total = args.batch * per;
    const uint2 *acc = reinterpret_cast<const uint2 *>(mac.acc);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t unit = i / per, r = i - unit * per;
        const uint32_t q = args.limb[(int)(r / n)].q;
        uint32_t s0 = 0, s1 = 0;
        for (int pt = 0; pt < parts; ++pt) {
            const uint2 a = __ldcs(acc + (unit * parts + pt) * per + r);
            s0 = ab_addmod(s0, a.x, q); s1 = ab_addmod(s1, a.y, q);
        }
        if (mac.add) {
            s0 = ab_addmod(s0, (uint32_t)__ldcs(mac.add + unit * mac.add_ustride + r), q);
            s1 = ab_addmod(s1, (uint32_t)__ldcs(mac.add + unit * mac.add_ustride + per + r), q);
        }
        if (mac.add0) s0 = ab_addmod(s0, (uint32_t)__ldcs(mac.add0 + unit * mac.add0_ustride + r), q);
        __stcs(mac.out + unit * mac.out_ustride + r, (int64_t)s0);
        __stcs(mac.out + unit * mac.out_ustride + per + r, (int64_t)s1);
    }
}

static size_t mac_smem(const ab_plan *p, int max_words, bool acc_smem) {
    const size_t nb = ((size_t)p->n + 3) & ~(size_t)3;
    return ((size_t)max_words + nb + (acc_smem ? 2 * (size_t)p->n : 0)) * sizeof(uint32_t);
}
// the running sums ride in shared memory only where four CTAs still fit an SM; otherwise in the caller's scratch
static bool mac_acc_in_smem(const ab_plan *p, int max_words) { return 4 * (mac_smem(p, max_words, true) + 1024) <= 228 * 1024; }
bool ab_mixed_mac_supported(const ab_plan *p, int32_t nlimbs) {
    if (ab_pow2_supported(p) || (p->n & 1)) return false;
    return ab_mixed_supported(p, AbView{nullptr, 1, 0, nlimbs, AB_LAYOUT_INTERLEAVED});
}
// a CTA of the fused inner product runs for milliseconds: when the batch alone gives fewer than about four waves of CTAs, cut the
// digit list of a unit into parts (each keeps at least eight digit transforms, so table loads and gathers stay amortised).
// Measured on the Tunnel.hs chain: 512 ciphertexts 34.7k -> 40.2k/s; at 4096 ciphertexts splitting does not pay (left at 1 part).
int32_t ab_mixed_mac_parts(const ab_plan *p, int64_t units, int32_t nlimbs, int32_t total_digits) {
    static const int force = getenv("AB_MX_MAC_PARTS") ? atoi(getenv("AB_MX_MAC_PARTS")) : 0;      // experiments only
    if (force > 0) return std::min(force, total_digits);
    if (units <= 0) return 1;
    const int64_t slots = (int64_t)p->sm_count * 5, want = 4 * slots;
    int64_t parts = (want + units * nlimbs - 1) / (units * nlimbs);
    parts = std::min<int64_t>(parts, total_digits / 8);
    // latency regime (round 2): a handful of ciphertexts -- the reference's own timed run is ONE evaluation (examples/HomomRLWR.hs:68) -- does
    // not even fill one wave with eight digits per CTA: go down to two digits per CTA as long as the grid stays within one wave
    // (measured on the Tunnel.hs chain, one ciphertext: see profiles/r02_small_batch_latency.jsonl)
    static const bool no_lat = getenv("AB_MX_MAC_NOLATENCY") && atoi(getenv("AB_MX_MAC_NOLATENCY")) != 0;      // A/B measurements only
    if (!no_lat && units * nlimbs * std::max<int64_t>(1, parts) < slots) {
        const int64_t fit = slots / (units * nlimbs);
        parts = std::max<int64_t>(parts, std::min<int64_t>(fit, total_digits / 2));
    }
    return (int32_t)std::max<int64_t>(1, parts);
}
int64_t ab_mixed_mac_scratch_words(const ab_plan *p, int64_t units, int32_t nlimbs) {
    const int words = std::max(p->mx_words_fwd[0], p->mx_words_fwd[1]);
    return mac_acc_in_smem(p, words) ? 0 : units * nlimbs * p->n;      // uint2 per position = one int64 word (per part when parts > 1)
}

void ab_launch_crt_mixed_mac(const ab_plan *p, int64_t units, int32_t nlimbs, const AbMxFused &fx, const AbMxMac &mac, cudaStream_t st) {
    AB_REQUIRE(ab_mixed_mac_supported(p, nlimbs), AB_ERR_UNSUPPORTED, "fused gadget inner product: plan not supported");
    AB_REQUIRE(fx.load_mode == AB_MX_LOAD_DIGITS && fx.nd >= 1 && mac.K >= 1 && mac.hm && mac.out, AB_ERR_ARG, "fused gadget inner product: bad arguments");
    if (units == 0) return;
    MxArgs args{};
    args.fx = fx;
    int max_words = 0;
    fill_mx_args(p, args, 0, nlimbs, false, max_words);
    args.batch = units * (mac.parts > 1 ? mac.parts : 1);      // CTAs per limb
    args.estride = 1; args.lstride = p->n; args.bstride = p->n * nlimbs;
    AbMxMac mac2 = mac;
    static const bool no_bits = getenv("AB_NO_BITS") && atoi(getenv("AB_NO_BITS")) != 0;      // A/B measurements only
    mac2.bits_p = no_bits ? 0 : p->mx_bits_p; mac2.bits_h = p->mx_bits_h;
    for (int l = 0; l < nlimbs; ++l) mac2.bits_lut[l] = (mac2.bits_p && (size_t)l < p->mx_bits_lut.size()) ? p->mx_bits_lut[l] : nullptr;
    const int parts = mac.parts > 1 ? mac.parts : 1;
    AB_REQUIRE(mac.n_out == 0 || (mac.rep_idx && (mac.n_out & 1) == 0), AB_ERR_ARG, "fused gadget inner product: replicated mode needs an index table and an even output size");
    const bool acc_smem = parts == 1 && mac.n_out == 0 && mac_acc_in_smem(p, std::max(p->mx_words_fwd[0], p->mx_words_fwd[1]));
    AB_REQUIRE(acc_smem || mac.acc, AB_ERR_ARG, "fused gadget inner product: scratch missing");
    // replicated mode: the transforms are small, so several of them fit shared memory side by side (about 40 KB of residues keeps
    // four to five CTAs per SM)
    const size_t nbw = (((size_t)p->n + 3) & ~(size_t)3) * sizeof(uint32_t);
    static const int force_g = getenv("AB_MX_MAC_GROUP") ? atoi(getenv("AB_MX_MAC_GROUP")) : 0;      // experiments only
    int group = 1;
    if (mac.n_out > 0 && !acc_smem) group = force_g > 0 ? std::min(force_g, 8) : (int)std::max<size_t>(1, std::min<size_t>(8, (40 * 1024) / nbw));
    mac2.group = group;
    const size_t smem = mac_smem(p, max_words, acc_smem) + (size_t)(group - 1) * nbw;
    static const int force_t = getenv("AB_MX_MAC_T") ? atoi(getenv("AB_MX_MAC_T")) : 0;      // experiments only
    const int T = force_t ? force_t : pick_threads(p->mx_fwd, (int)p->n, smem);
    AB_REQUIRE(units * parts * nlimbs < (1ll << 31), AB_ERR_ARG, "batch too large for one launch");
    const dim3 grid((unsigned)(units * parts * nlimbs));
    if (acc_smem) ab_mx_launch_mac_smem(grid, T, smem, st, args, mac2);
    else ab_mx_launch_mac_glob(grid, T, smem, st, args, mac2);
    ab_count_launch();
    ab_check_launch("crt_mixed_mac");
    if (parts > 1) {
        const int64_t work = units * nlimbs * (mac.n_out > 0 ? mac.n_out : p->n);
        const unsigned g = (unsigned)std::max<int64_t>(1, std::min<int64_t>((work + 255) / 256, (int64_t)p->sm_count * 16));
        args.batch = units;      // the combine pass walks over units, not CTAs
        k_mac_combine<<<g, 256, 0, st>>>(args, mac2);
        ab_count_launch();
        ab_check_launch("mac_combine");
    }
}

