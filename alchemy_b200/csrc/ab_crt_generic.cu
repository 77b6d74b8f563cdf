// ab_crt_generic.cu -- batched CRT / crtInv for arbitrary cyclotomic index m (mixed radix).
//
// Replaces lol-cpp `tensorCRTRq` / `tensorCRTInvRq` [RECALL] for the non-power-of-two rings of
// examples/Common.hs:41-54 (m' = 11648 ... 54600) and BASELINE config 3 (m = 40320).
// CRT_m is the Kronecker product over prime powers of
//     CRT_{p^e} = (DFT_{p^(e-1)} (x) I_{p-1}) * That * (I_{p^(e-1)} (x) CRT_p)
// (Lol paper App. C); ab_plan.cpp compiles it into a short program of steps, each a set of
// independent small dense mat-vecs (CRT_p: (p-1)x(p-1), DFT_p: pxp) or radix-2 butterflies on
// strided elements.  One CTA owns one (ring element, limb): the n = phi(m) residues are read once
// from HBM (int64 -> uint32), every step runs out of shared memory, and the result is written
// once -- 16 n bytes of HBM traffic per transform regardless of the number of stages.
#include "ab_internal.h"
#include "ab_modarith.cuh"

struct GenArgs {
    struct {
        uint32_t q;
        const uint2 *tab;
    } limb[AB_MAX_LIMBS];
    int32_t nlimbs, nsteps, n;
    int64_t ntrans;
    int64_t estride, lstride, bstride;   // int64 strides of coefficient, limb, batch element
    const AbGenStep *steps;
};

template <int P>
__device__ __forceinline__ void matvec_item(uint32_t *sm, const AbGenStep &s, int v, const uint2 *__restrict__ tab, uint32_t q) {
    const int tb = v / s.estride, off = v - tb * s.estride;
    uint32_t *e = sm + tb * P * s.estride + off;
    const int row = (s.kind == 0) ? (tb % s.h) : (off / s.sdiv);
    const uint2 *sc = tab + s.sc_off + row * P;
    const uint2 *M = tab + s.mat_off;
    uint32_t x[P], o[P];
#pragma unroll
    for (int k = 0; k < P; ++k) x[k] = e[k * s.estride];
    if (s.mode == 1) {
#pragma unroll
        for (int k = 0; k < P; ++k) x[k] = ab_shoup(x[k], __ldg(sc + k), q);
    }
#pragma unroll
    for (int r = 0; r < P; ++r) {
        uint32_t acc = 0;
#pragma unroll
        for (int c = 0; c < P; ++c) acc = ab_addmod(acc, ab_shoup(x[c], __ldg(M + r * P + c), q), q);
        o[r] = acc;
    }
    if (s.mode == 2) {
#pragma unroll
        for (int k = 0; k < P; ++k) o[k] = ab_shoup(o[k], __ldg(sc + k), q);
    }
#pragma unroll
    for (int k = 0; k < P; ++k) e[k * s.estride] = o[k];
}

// any P up to 64 (primes up to 61): same arithmetic with runtime loops
__device__ void matvec_item_rt(uint32_t *sm, const AbGenStep &s, int v, const uint2 *__restrict__ tab, uint32_t q) {
    const int P = s.P;
    const int tb = v / s.estride, off = v - tb * s.estride;
    uint32_t *e = sm + tb * P * s.estride + off;
    const int row = (s.kind == 0) ? (tb % s.h) : (off / s.sdiv);
    const uint2 *sc = tab + s.sc_off + row * P;
    const uint2 *M = tab + s.mat_off;
    uint32_t x[64], o[64];
    for (int k = 0; k < P; ++k) {
        x[k] = e[k * s.estride];
        if (s.mode == 1) x[k] = ab_shoup(x[k], __ldg(sc + k), q);
    }
    for (int r = 0; r < P; ++r) {
        uint32_t acc = 0;
        for (int c = 0; c < P; ++c) acc = ab_addmod(acc, ab_shoup(x[c], __ldg(M + r * P + c), q), q);
        o[r] = (s.mode == 2) ? ab_shoup(acc, __ldg(sc + r), q) : acc;
    }
    for (int k = 0; k < P; ++k) e[k * s.estride] = o[k];
}

__device__ __forceinline__ void radix2_item(uint32_t *sm, const AbGenStep &s, int v, const uint2 *__restrict__ tab, uint32_t q) {
    const int tb = v / s.estride, off = v - tb * s.estride;
    uint32_t *e = sm + tb * 2 * s.estride + off;
    const uint2 w = __ldg(tab + s.sc_off + off / s.sdiv);
    uint32_t a = e[0], b = e[s.estride];
    if (s.mode == 1) {   // forward: twiddle the upper input, then butterfly
        b = ab_shoup(b, w, q);
        e[0] = ab_addmod(a, b, q);
        e[s.estride] = ab_submod(a, b, q);
    } else {             // inverse: butterfly, then twiddle the lower output
        uint32_t sum = ab_addmod(a, b, q), d = ab_submod(a, b, q);
        if (s.flags & 1) sum = ab_shoup(sum, __ldg(tab + s.sc_off + s.h), q);
        e[0] = sum;
        e[s.estride] = ab_shoup(d, w, q);
    }
}

__global__ void __launch_bounds__(512) k_crt_generic(const GenArgs args, int64_t *__restrict__ y) {
    extern __shared__ uint32_t sm[];
    const int64_t trans = blockIdx.x;
    const int limb = (int)(trans % args.nlimbs);
    const int64_t b = trans / args.nlimbs;
    const uint32_t q = args.limb[limb].q;
    const uint2 *__restrict__ tab = args.limb[limb].tab;
    int64_t *yp = y + b * args.bstride + limb * args.lstride;
    const int n = args.n;
    for (int j = threadIdx.x; j < n; j += blockDim.x) sm[j] = (uint32_t)yp[j * args.estride];
    __syncthreads();
    for (int si = 0; si < args.nsteps; ++si) {
        const AbGenStep s = args.steps[si];
        for (int v = threadIdx.x; v < s.nwork; v += blockDim.x) {
            if (s.kind == 1) radix2_item(sm, s, v, tab, q);
            else switch (s.P) {
                case 2: matvec_item<2>(sm, s, v, tab, q); break;
                case 3: matvec_item<3>(sm, s, v, tab, q); break;
                case 4: matvec_item<4>(sm, s, v, tab, q); break;
                case 5: matvec_item<5>(sm, s, v, tab, q); break;
                case 6: matvec_item<6>(sm, s, v, tab, q); break;
                case 7: matvec_item<7>(sm, s, v, tab, q); break;
                case 12: matvec_item<12>(sm, s, v, tab, q); break;
                case 13: matvec_item<13>(sm, s, v, tab, q); break;
                default: matvec_item_rt(sm, s, v, tab, q); break;
            }
        }
        __syncthreads();
    }
    for (int j = threadIdx.x; j < n; j += blockDim.x) yp[j * args.estride] = (int64_t)sm[j];
}

void ab_launch_crt_generic(const ab_plan *p, const AbView &v, bool inverse, cudaStream_t st) {
    AB_REQUIRE(p->has_crt, AB_ERR_NO_ROOT, "plan has no CRT tables");
    if (v.batch == 0) return;
    for (auto &a : p->axes) AB_REQUIRE(a.p <= 61, AB_ERR_UNSUPPORTED, "prime factors above 61 are not supported");
    GenArgs args{};
    args.nlimbs = v.nlimbs;
    args.n = (int32_t)p->n;
    args.ntrans = v.batch * v.nlimbs;
    const auto &steps = inverse ? p->steps_inv : p->steps_fwd;
    args.nsteps = (int32_t)steps.size();
    args.steps = inverse ? p->d_steps_inv : p->d_steps_fwd;
    if (v.layout == AB_LAYOUT_PLANAR) { args.estride = 1; args.lstride = p->n; args.bstride = p->n * v.nlimbs; }
    else { args.estride = v.nlimbs; args.lstride = 1; args.bstride = p->n * v.nlimbs; }
    for (int l = 0; l < v.nlimbs; ++l) {
        const AbLimbDev &ld = p->limb[v.limb0 + l];
        args.limb[l].q = ld.q;
        args.limb[l].tab = reinterpret_cast<const uint2 *>(inverse ? ld.gen_inv : ld.gen_fwd);
    }
    const size_t smem = (size_t)p->n * sizeof(uint32_t);
    AB_REQUIRE(smem <= 227 * 1024, AB_ERR_UNSUPPORTED, "ring too large for the shared-memory transform");
    static thread_local size_t smem_set = 0;
    if (smem > 48 * 1024 && smem > smem_set) {
        AB_CHECK_CUDA(cudaFuncSetAttribute(k_crt_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        smem_set = 227 * 1024;
    }
    int threads = p->n >= 8192 ? 512 : (p->n >= 2048 ? 256 : 128);
    k_crt_generic<<<(unsigned)args.ntrans, threads, smem, st>>>(args, v.y);
    ab_count_launch();
    ab_check_launch("crt_generic");
}
