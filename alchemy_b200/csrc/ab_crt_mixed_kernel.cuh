// ab_crt_mixed_kernel.cuh -- the transform kernel k_crt_mixed<INV> and its launch wrapper.  Included by two translation
// units (ab_crt_mixed.cu: forward, ab_crt_mixed_inv.cu: inverse) with AB_MX_KERNEL_INV set, so the two instantiations compile in parallel.
#pragma once
#include "ab_crt_mixed_dev.cuh"

template <bool INV>
__global__ void __launch_bounds__(256, AB_MX_MINB) k_crt_mixed(const __grid_constant__ MxArgs args, int64_t *__restrict__ y) {
    extern __shared__ __align__(16) uint32_t smem[];
    // one-dimensional grid over (ring element, limb), limb-major; AB_MX_INTERLEAVE=1 interleaves the limbs (block b = ring element
    // b / L, limb b % L) so that the ALU-heavy transforms of the moduli above 2^30 sit beside the IMAD-heavy ones of the narrow moduli
    const int limb = args.interleave ? (int)(blockIdx.x % (unsigned)args.nlimbs) : (int)(blockIdx.x / (unsigned)args.batch);
    const unsigned bx = args.interleave ? blockIdx.x / (unsigned)args.nlimbs : blockIdx.x % (unsigned)args.batch;
    const bool wide = args.limb[limb].wide != 0;
    const int tab_words = args.tab_words[wide ? 1 : 0];
    uint32_t *tab = smem;                          // [tab_words]
    uint32_t *sm = smem + tab_words;               // [n]
    const uint32_t q = args.limb[limb].q, qinv_neg = args.limb[limb].qinv_neg;
    const int n = args.n;
    int64_t *yp = y + bx * args.bstride + limb * args.lstride;
    {   // constant tables of this limb (L2-resident, a few KB)
        const uint4 *src = reinterpret_cast<const uint4 *>(args.limb[limb].tab);
        uint4 *dst = reinterpret_cast<uint4 *>(tab);
        for (int i = threadIdx.x; i < tab_words / 4; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    bool skip_program = false;
    if (args.fx.load_mode == AB_MX_LOAD_DIGITS) {
        // the residues of this transform are gadget digit d of a powerful-basis element (optionally gathered
        // through an index table: coeffs . embed of the tunnel), centred and reduced into this CTA's modulus
        const AbMxFused &fx = args.fx;
        const int d = (int)(bx % fx.nd);
        const int64_t unit = bx / fx.nd;
        if (mx_digit_is_trivial(fx, d, limb)) {
            // TrivGad digit of limb l reduced mod q_l is the element itself: its CRT is the caller's CRT-basis input
            const int4 *src = reinterpret_cast<const int4 *>(fx.triv_src + unit * fx.triv_ustride + (int64_t)limb * n);
            int4 *dst = reinterpret_cast<int4 *>(yp);
            for (int i = threadIdx.x; i < n / 2; i += blockDim.x) __stcs(dst + i, __ldcs(src + i));
            return;
        }
        mx_load_digit(fx, d, fx.gK > 1 ? unit / fx.gK : unit, fx.gK > 1 ? (int)(unit % fx.gK) : 0, n, q, sm);
    } else if (args.fx.load_mode == AB_MX_LOAD_GARNER) {
        const AbMxFused &fx = args.fx;
        const int k = fx.gar_k;
        const int64_t *src = fx.src + bx * fx.src_ustride;
        uint32_t qmod[AB_MX_MAX_DROP];                 // q_i mod q of this CTA
#pragma unroll
        for (int i = 0; i < AB_MX_MAX_DROP; ++i) qmod[i] = i < k ? fx.gar_q[i] % q : 0u;
        const uint64_t mu = ~0ull / q;                 // Barrett constant of this CTA's modulus (floor((2^64 - 1) / q) = floor(2^64 / q), q odd > 1)
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            uint32_t v[AB_MX_MAX_DROP];
            int32_t r[AB_MX_MAX_DROP];
#pragma unroll
            for (int i = 0; i < AB_MX_MAX_DROP; ++i) v[i] = i < k ? (uint32_t)__ldcs(src - (int64_t)i * n + j) : 0u;   // drop order: last limb first
#pragma unroll
            for (int i = 0; i < AB_MX_MAX_DROP; ++i) {
                if (i >= k) break;
                r[i] = ab_center(v[i], fx.gar_q[i]);
#pragma unroll
                for (int t = i + 1; t < AB_MX_MAX_DROP; ++t) {
                    if (t >= k) break;
                    const uint32_t qt = fx.gar_q[t];
                    v[t] = ab_shoup(ab_submod(v[t], mx_reduce_signed(r[i], qt), qt), fx.gar_inv[i][t], fx.gar_inv_s[i][t], qt);
                }
            }
            uint32_t acc = mx_reduce_signed(r[k - 1], q);
            for (int i = k - 2; i >= 0; --i) acc = ab_addmod(ab_mulmod(acc, qmod[i], q, mu), mx_reduce_signed(r[i], q), q);
            sm[j] = acc;
        }
    } else if (args.fx.load_mode == AB_MX_LOAD_GATHER) {
        const AbMxFused &fx = args.fx;
        const int64_t su = bx / fx.gK;
        const int32_t *__restrict__ gt = fx.gather + (int64_t)(bx % fx.gK) * n;
        const int64_t *__restrict__ src = fx.src + su * fx.src_ustride + (int64_t)limb * fx.n_src;
        for (int j0 = threadIdx.x; j0 < n; j0 += 4 * blockDim.x) {
            int32_t g[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int j = j0 + u * blockDim.x; g[u] = j < n ? __ldg(gt + j) : -1; }
            uint32_t v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = g[u] >= 0 ? (uint32_t)__ldg(src + g[u]) : 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int j = j0 + u * blockDim.x; if (j < n) sm[j] = v[u]; }
        }
    } else if (args.fx.load_mode == AB_MX_LOAD_SRC) {
        // out of place: read the transform from another planar buffer (e.g. c2 of a degree-2 ciphertext)
        const int4 *src = reinterpret_cast<const int4 *>(args.fx.src + bx * args.fx.src_ustride + (int64_t)(args.fx.src_limb0 + limb) * n);
        uint2 *dst = reinterpret_cast<uint2 *>(sm);
        for (int i = threadIdx.x; i < n / 2; i += blockDim.x) {
            const int4 v = __ldcs(src + i);
            dst[i] = make_uint2((uint32_t)v.x, (uint32_t)v.z);
        }
    } else if (args.estride == 1) {
        const int4 *src = reinterpret_cast<const int4 *>(yp);
        uint2 *dst = reinterpret_cast<uint2 *>(sm);
        for (int i = threadIdx.x; i < n / 2; i += blockDim.x) {
            const int4 v = __ldcs(src + i);
            dst[i] = make_uint2((uint32_t)v.x, (uint32_t)v.z);
        }
    } else {
        for (int j = threadIdx.x; j < n; j += blockDim.x) sm[j] = (uint32_t)yp[j * args.estride];
    }
    (void)skip_program;
    __syncthreads();
    const bool l_on = args.fx.l_period <= 1 || ((args.fx.l_mask >> ((bx + args.fx.l_phase) % args.fx.l_period)) & 1) != 0;
    if (!INV && args.fx.pre_l && l_on) mx_l_pass(sm, n, args.fx, q, false);
    if (wide) run_program<INV, true>(args, sm, tab, q, qinv_neg);
    else run_program<INV, false>(args, sm, tab, q, qinv_neg);
    if (INV && args.fx.post_linv && l_on) mx_l_pass(sm, n, args.fx, q, true);
    if (args.fx.store_mode == AB_MX_STORE_SUBSCALE) {
        // RNS rescale epilogue: out = (in - x) * q_drop^-1 mod q   (x = CRT of the lifted dropped limb)
        const int64_t *in = args.fx.sub_in + bx * args.fx.sub_ustride + (int64_t)limb * n;
        const uint32_t sc = args.fx.scale[limb], scs = args.fx.scale_s[limb];
        const int4 *src = reinterpret_cast<const int4 *>(in);
        int4 *dst = reinterpret_cast<int4 *>(yp);
        const uint2 *xs = reinterpret_cast<const uint2 *>(sm);
        for (int i = threadIdx.x; i < n / 2; i += blockDim.x) {
            const int4 a = __ldcs(src + i);
            const uint2 x = xs[i];
            const uint32_t r0 = ab_shoup(ab_submod((uint32_t)a.x, x.x, q), sc, scs, q);
            const uint32_t r1 = ab_shoup(ab_submod((uint32_t)a.z, x.y, q), sc, scs, q);
            __stcs(dst + i, make_int4((int)r0, 0, (int)r1, 0));
        }
    } else if (args.estride == 1) {
        int4 *dst = reinterpret_cast<int4 *>(yp);
        const uint2 *src = reinterpret_cast<const uint2 *>(sm);
        for (int i = threadIdx.x; i < n / 2; i += blockDim.x) {
            const uint2 v = src[i];
            __stcs(dst + i, make_int4((int)v.x, 0, (int)v.y, 0));
        }
    } else {
        for (int j = threadIdx.x; j < n; j += blockDim.x) yp[j * args.estride] = (int64_t)sm[j];
    }
}

#if AB_MX_KERNEL_INV
void ab_mx_launch_inv(dim3 grid, int T, size_t smem, cudaStream_t st, const MxArgs &args, int64_t *y) {
#else
void ab_mx_launch_fwd(dim3 grid, int T, size_t smem, cudaStream_t st, const MxArgs &args, int64_t *y) {
#endif
    static thread_local bool ready = false;
    if (!ready) {
        AB_CHECK_CUDA(cudaFuncSetAttribute(k_crt_mixed<AB_MX_KERNEL_INV != 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        ready = true;
    }
    k_crt_mixed<AB_MX_KERNEL_INV != 0><<<grid, T, smem, st>>>(args, y);
}
