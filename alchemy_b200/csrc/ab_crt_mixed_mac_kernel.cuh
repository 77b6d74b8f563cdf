// ab_crt_mixed_mac_kernel.cuh -- the fused gadget inner product kernel and its launch wrapper.  Included by two translation
// units with AB_MX_MAC_SMEM set (running sums in shared memory or in the caller's scratch), so they compile in parallel.
#pragma once
#include "ab_crt_mixed_dev.cuh"

// ---------------------------------------------------------------------------------------------
// Fused gadget inner product: out_c = add_c + sum_{k,d} CRT(reduce(digit_{k,d}(src))) * hint_{k,d,c},  c in {0,1}.
// The core of keySwitchQuadCirc (Eval.hs:133) and of tunnel (Eval.hs:134).  One CTA owns (unit, limb) and keeps
// the residues of the digit being transformed (uint32 [n]) and the two running sums (uint2 [n], fully reduced)
// in shared memory for the whole loop: the K * nd digit transforms never touch HBM.  Per digit: load + centre +
// digit + reduce (source and gather table from L2), the forward program, then one pass x * (h0, h1) with a
// single-term Montgomery reduction each (T = x h 2^32 < 2^62, T + m q < 2^64 for q < 2^31).
// ACC_SMEM = false: the running sums live in a global scratch slice owned by this CTA (thread t always touches the
// same positions j = t, t + T, ..., so no synchronisation is needed; with one wave of CTAs resident the slices stay
// in L2) and shared memory holds only the residues: the 4 x 256-thread residency of the plain transform is kept.
template <int MAXT, int MINB, bool ACC_SMEM>
__global__ void __launch_bounds__(MAXT, MINB) k_crt_mixed_mac(const __grid_constant__ MxArgs args, const __grid_constant__ AbMxMac mac) {
    extern __shared__ __align__(16) uint32_t smem[];
    const int limb = args.interleave ? (int)(blockIdx.x % (unsigned)args.nlimbs) : (int)(blockIdx.x / (unsigned)args.batch);      // as in k_crt_mixed
    const unsigned bx = args.interleave ? blockIdx.x / (unsigned)args.nlimbs : blockIdx.x % (unsigned)args.batch;
    const bool wide = args.limb[limb].wide != 0;
    const int tab_words = args.tab_words[wide ? 1 : 0];
    const int n = args.n, nb = (n + 3) & ~3;
    uint32_t *tab = smem;
    uint32_t *sm = smem + tab_words;
    const uint32_t q = args.limb[limb].q, qinv_neg = args.limb[limb].qinv_neg;
    const AbMxFused &fx = args.fx;
    const int parts = mac.parts > 1 ? mac.parts : 1;
    const int64_t unit = bx / parts, slice = bx;          // slice = unit * parts + part
    const int part = (int)(bx - unit * parts);
    const int nm = (!ACC_SMEM && mac.n_out > 0) ? mac.n_out : n;       // positions of the inner product (hint, sums, output)
    const int32_t *rep = (!ACC_SMEM && mac.n_out > 0) ? mac.rep_idx : nullptr;
    uint2 *acc = ACC_SMEM ? reinterpret_cast<uint2 *>(sm + nb)
                          : reinterpret_cast<uint2 *>(mac.acc) + (slice * args.nlimbs + limb) * (int64_t)nm;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(args.limb[limb].tab);
        uint4 *dst = reinterpret_cast<uint4 *>(tab);
        for (int i = threadIdx.x; i < tab_words / 4; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    uint32_t *tbuf = (mac.tbuf && fx.gbase != 0) ? reinterpret_cast<uint32_t *>(mac.tbuf) + (slice * args.nlimbs + limb) * (int64_t)nb : nullptr;
    // subset-sum tables apply to base 2 with the digits cut from the gathered residues, on a 16-byte aligned residue buffer
    const uint32_t *bits_lut = (tbuf && fx.gbase == 2 && mac.bits_p > 0 && (n % mac.bits_p) == 0 && ((uintptr_t)sm & 15) == 0) ? mac.bits_lut[limb] : nullptr;
    const uint2 *hm = reinterpret_cast<const uint2 *>(mac.hm) + (int64_t)limb * nm;
    const int64_t hstride = (int64_t)args.nlimbs * nm;
    bool first = true;
    const int total_digits = mac.K * fx.nd, per_part = (total_digits + parts - 1) / parts;
    const int kd0 = part * per_part, kd1 = min(total_digits, kd0 + per_part);
    // G digit transforms are kept side by side in shared memory (replicated mode: they are small) and ONE inner-product pass
    // consumes all of them: the running sums, the index table and the reductions are touched once per G digits, and the G
    // products per output are accumulated exactly in 64 bits (narrow moduli: terms below 2^60, G <= 8) before one Montgomery step
    const int G = ACC_SMEM ? 1 : max(1, min(mac.group, 8));
    const int lazy_terms = args.limb[limb].lazy;      // 8 for q < 2^30; 2 ... 4 above (round 1 reduced every term of a wide limb separately)
    for (int kdg = kd0; kdg < kd1; kdg += G) {
        const int gcount = min(G, kd1 - kdg);
        for (int gi = 0; gi < gcount; ++gi) {
            const int kd = kdg + gi;
            const int k = kd / fx.nd, d = kd - k * fx.nd;
            uint32_t *buf = sm + gi * nb;
            const bool triv = mx_digit_is_trivial(fx, d, limb);
            int first_step = 0;
            if (triv) {
                const int4 *src = reinterpret_cast<const int4 *>(fx.triv_src + unit * fx.triv_ustride + (int64_t)limb * n);
                uint2 *dst = reinterpret_cast<uint2 *>(buf);
                for (int i = threadIdx.x; i < n / 2; i += blockDim.x) {
                    const int4 v = __ldcs(src + i);
                    dst[i] = make_uint2((uint32_t)v.x, (uint32_t)v.z);
                }
            } else if (tbuf) {
                if (kd == kd0 || d == 0 || fx.dig[d].limb != fx.dig[d - 1].limb) {
                    __syncthreads();      // every earlier reader of t is done
                    mx_gather_centered(fx, d, unit, k, n, tbuf);
                    __syncthreads();      // the digit pass reads t with a different thread-to-position map
                }
                // binary gadget: digit and first transform step in one pass (subset-sum tables), except for the last digit
                if (bits_lut && !fx.dig[d].last) first_step = mx_bits_dispatch(mac.bits_p, fx, d, n, q, tbuf, bits_lut, buf) ? 1 : 0;
                if (!first_step) mx_digit_from_t(fx, d, n, q, tbuf, buf);
            } else {
                mx_load_digit(fx, d, unit, k, n, q, buf);
            }
            __syncthreads();
            if (!triv) {
                if (wide) run_program<false, true>(args, buf, tab, q, qinv_neg, first_step);
                else run_program<false, false>(args, buf, tab, q, qinv_neg, first_step);
            }
        }
        const uint2 *hp = hm + (int64_t)kdg * hstride;            // hint of digit kdg + gi at hp + gi * hstride
        auto fold = [&](uint64_t &s, uint32_t &a, bool many) {    // a += REDC(s), s = 0; a single term reduces to below 2q
            uint32_t t = (uint32_t)((s + (uint64_t)((uint32_t)s * qinv_neg) * q) >> 32);
            if (many) t = ab_csub(t, 2 * q);
            a = ab_addmod(ab_csub(t, q), a, q);
            s = 0;
        };
        // four consecutive positions per thread and step: 16-byte loads of the hint pairs, the running sums and the
        // residues; thread t owns the same positions for every digit (the sums need no synchronisation)
        const int T = blockDim.x, n4 = nm >> 2;
        uint4 *a4 = reinterpret_cast<uint4 *>(acc);
        const int4 *r4 = reinterpret_cast<const int4 *>(rep);
        for (int g = threadIdx.x; g < n4; g += T) {
            uint4 aa = make_uint4(0, 0, 0, 0), ab2 = aa;
            if (!first) {
                if (ACC_SMEM) { aa = a4[2 * g]; ab2 = a4[2 * g + 1]; }
                else { aa = __ldcg(a4 + 2 * g); ab2 = __ldcg(a4 + 2 * g + 1); }
            }
            int4 ix = make_int4(4 * g, 4 * g + 1, 4 * g + 2, 4 * g + 3);
            if (rep) ix = __ldg(r4 + g);
            uint64_t s00 = 0, s01 = 0, s10 = 0, s11 = 0, s20 = 0, s21 = 0, s30 = 0, s31 = 0;      // [position][component]
            // running pointers instead of index arithmetic per term: the hint pair of this group in digit gi, and the residues
            const char *hb = reinterpret_cast<const char *>(hp) + (size_t)g * 32;
            const size_t hstep = (size_t)hstride * sizeof(uint2);
            const uint32_t *b0 = sm + (rep ? ix.x : 4 * g), *b1 = sm + ix.y, *b2 = sm + ix.z, *b3 = sm + ix.w;
            auto term = [&](int) {
                const uint4 ha = __ldg(reinterpret_cast<const uint4 *>(hb)), hb2 = __ldg(reinterpret_cast<const uint4 *>(hb + 16));
                uint4 x;
                if (rep && !mac.rep4) x = make_uint4(*b0, *b1, *b2, *b3);
                else x = *reinterpret_cast<const uint4 *>(b0);
                hb += hstep; b0 += nb; b1 += nb; b2 += nb; b3 += nb;
                s00 += (uint64_t)x.x * ha.x; s01 += (uint64_t)x.x * ha.y;
                s10 += (uint64_t)x.y * ha.z; s11 += (uint64_t)x.y * ha.w;
                s20 += (uint64_t)x.z * hb2.x; s21 += (uint64_t)x.z * hb2.y;
                s30 += (uint64_t)x.w * hb2.z; s31 += (uint64_t)x.w * hb2.w;
            };
            auto fold_all = [&](bool many) {
                fold(s00, aa.x, many); fold(s01, aa.y, many); fold(s10, aa.z, many); fold(s11, aa.w, many);
                fold(s20, ab2.x, many); fold(s21, ab2.y, many); fold(s30, ab2.z, many); fold(s31, ab2.w, many);
            };
            if (gcount == 1) {          // the common plain case: no loop, constant offsets
                term(0);
                fold_all(false);
            } else {
                int terms = 0;
                for (int gi = 0; gi < gcount; ++gi) {
                    term(gi);
                    if (++terms == lazy_terms || gi + 1 == gcount) { fold_all(terms > 1); terms = 0; }
                }
            }
            if (ACC_SMEM) { a4[2 * g] = aa; a4[2 * g + 1] = ab2; }
            else { __stcg(a4 + 2 * g, aa); __stcg(a4 + 2 * g + 1, ab2); }
        }
        for (int j = 4 * n4 + threadIdx.x; j < nm; j += T) {
            uint2 a = make_uint2(0, 0);
            if (!first) a = ACC_SMEM ? acc[j] : __ldcg(acc + j);
            const int jx = rep ? __ldg(rep + j) : j;
            for (int gi = 0; gi < gcount; ++gi) {
                const uint2 h = __ldg(hp + (int64_t)gi * hstride + j);
                const uint32_t x = sm[gi * nb + jx];
                uint64_t s0 = (uint64_t)x * h.x, s1 = (uint64_t)x * h.y;
                fold(s0, a.x, false); fold(s1, a.y, false);
            }
            if (ACC_SMEM) acc[j] = a;
            else __stcg(acc + j, a);
        }
        first = false;
        __syncthreads();       // the next group overwrites the residue buffers
    }
    if (first && !ACC_SMEM) {       // a part that received no digit (parts does not divide the digit count): its partial sum is zero
        for (int j = threadIdx.x; j < nm; j += blockDim.x) __stcg(acc + j, make_uint2(0, 0));
    }
    if (parts == 1) {
        const int64_t per = hstride;
        int64_t *o0 = mac.out + unit * mac.out_ustride + (int64_t)limb * nm, *o1 = o0 + per;
        const int64_t *a0 = mac.add ? mac.add + unit * mac.add_ustride + (int64_t)limb * nm : nullptr;
        const int64_t *a1 = a0 ? a0 + per : nullptr;
        const int64_t *l0 = mac.add0 ? mac.add0 + unit * mac.add0_ustride + (int64_t)limb * nm : nullptr;
        for (int i = threadIdx.x; i < nm / 2; i += blockDim.x) {
            // positions 2i, 2i+1 were accumulated by other threads of this CTA: the barrier after the last digit orders them
            const uint2 u = ACC_SMEM ? acc[2 * i] : __ldcg(acc + 2 * i), v = ACC_SMEM ? acc[2 * i + 1] : __ldcg(acc + 2 * i + 1);
            uint32_t r00 = u.x, r01 = v.x, r10 = u.y, r11 = v.y;      // r[c][column]
            if (a0) {
                const int4 x0 = __ldcs(reinterpret_cast<const int4 *>(a0) + i), x1 = __ldcs(reinterpret_cast<const int4 *>(a1) + i);
                r00 = ab_addmod(r00, (uint32_t)x0.x, q); r01 = ab_addmod(r01, (uint32_t)x0.z, q);
                r10 = ab_addmod(r10, (uint32_t)x1.x, q); r11 = ab_addmod(r11, (uint32_t)x1.z, q);
            }
            if (l0) {
                const int4 x0 = __ldcs(reinterpret_cast<const int4 *>(l0) + i);
                r00 = ab_addmod(r00, (uint32_t)x0.x, q); r01 = ab_addmod(r01, (uint32_t)x0.z, q);
            }
            __stcs(reinterpret_cast<int4 *>(o0) + i, make_int4((int)r00, 0, (int)r01, 0));
            __stcs(reinterpret_cast<int4 *>(o1) + i, make_int4((int)r10, 0, (int)r11, 0));
        }
    }
}

#if AB_MX_MAC_SMEM
void ab_mx_launch_mac_smem(dim3 grid, int T, size_t smem, cudaStream_t st, const MxArgs &args, const AbMxMac &mac) {
#else
void ab_mx_launch_mac_glob(dim3 grid, int T, size_t smem, cudaStream_t st, const MxArgs &args, const AbMxMac &mac) {
#endif
    static thread_local bool ready = false;
    if (!ready) {
        AB_CHECK_CUDA(cudaFuncSetAttribute(k_crt_mixed_mac<256, AB_MX_MAC_MINB, AB_MX_MAC_SMEM != 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        ready = true;
    }
    k_crt_mixed_mac<256, AB_MX_MAC_MINB, AB_MX_MAC_SMEM != 0><<<grid, T, smem, st>>>(args, mac);
}
