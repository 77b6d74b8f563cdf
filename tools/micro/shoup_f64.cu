// Micro-benchmark: the Harvey butterfly with its Shoup quotient taken (a) by IMAD.HI, (b) by the FP64 pipe -- sm_100a.
// Question it answers: IMAD.HI / IMAD.WIDE issue at about a third of the rate of a 32-bit IMAD on the FMA-heavy pipe; can the quotient
// floor(y w / q) of a twiddle multiplication come from one DFMA (FP64 pipe, runs beside the integer pipe) instead?
//   y52 = the double 2^52 + y, made by pairing y with the high word 0x43300000 (no conversion instruction)
//   R   = fma(y52, winv, C),  winv = floor(w 2^52 / q) / 2^52,  C = 2^52 + 2^51 - 1 - 2^52 winv   (both exact doubles)
//       = round(y winv) + 2^52 + 2^51 - 1 with ONE rounding, so the low word of R is k - 1, k = round(y w / q) (off by one only within 2^-20 of a tie)
//   t   = y w - (k - 1) q  in (0.49 q, 1.51 q): inside the [0, 2q) the lazy butterfly needs
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o shoup_f64 shoup_f64.cu ; run: ./shoup_f64
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t csub(uint32_t r, uint32_t q) { return min(r, r - q); }

struct Tw { uint32_t w, ws; double winv, c; };

template <int MODE> __device__ __forceinline__ uint32_t mul_lazy(uint32_t y, const Tw &t, uint32_t q) {
    if (MODE == 0) return y * t.w - __umulhi(y, t.ws) * q;                                  // [0, 2q)
    if (MODE == 1) {                                                                       // DADD + DFMA, no per-twiddle C
        const double yd = __hiloint2double(0x43300000, (int)y) - 4503599627370496.0;
        const double R = fma(yd, t.winv, 6755399441055743.0);                              // 2^52 + 2^51 - 1
        return y * t.w - (uint32_t)__double2loint(R) * q;
    }
    const double R = fma(__hiloint2double(0x43300000, (int)y), t.winv, t.c);               // one DFMA
    return y * t.w - (uint32_t)__double2loint(R) * q;
}
template <int MODE> __device__ __forceinline__ void bf(uint32_t &X, uint32_t &Y, const Tw &t, uint32_t q) {
    const uint32_t x = csub(X, 2 * q), m = mul_lazy<MODE>(Y, t, q);
    X = x + m; Y = x - m + 2 * q;
}
// MODE 3: IMAD.HI alone; MODE 4: 32-bit IMAD alone (same loop shape)
template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t *out, int iters, uint32_t q, const Tw *tw) {
    uint32_t x[16];
    const uint32_t seed = (blockIdx.x * 256 + threadIdx.x) * 2654435761u + 12345u;
    for (int i = 0; i < 16; ++i) x[i] = (seed + i * 977u) % q;
    if (MODE >= 3) {
        const uint32_t a = tw[threadIdx.x & 7].ws;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int s = 0; s < 4; ++s)
#pragma unroll
                for (int i = 0; i < 16; ++i) x[i] = (MODE == 3 ? __umulhi(x[(i + 1 + s) & 15], a) : x[(i + 1 + s) & 15] * a) + x[i];
    } else {
        Tw t[4];
        for (int s = 0; s < 4; ++s) t[s] = tw[(threadIdx.x + s) & 7];
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int s = 0; s < 4; ++s)
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (!(i & (1 << s))) bf<MODE>(x[i], x[i | (1 << s)], t[s], q);
    }
    uint32_t r = 0;
    for (int i = 0; i < 16; ++i) r += MODE >= 3 ? x[i] : (x[i] % q) * (i + 1);
    out[blockIdx.x * 256 + threadIdx.x] = r;
}

template <int MODE>
double run(const char *name, uint32_t *out, int sms, uint32_t q, const Tw *tw, uint32_t *host) {
    const int iters = 2048, blocks = sms * 4;
    k<MODE><<<blocks, 256>>>(out, 4, q, tw);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, iters, q, tw);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)blocks * 256 * iters * (MODE >= 3 ? 64 : 32);      // butterflies (or multiplies) per thread and iteration
    printf("%-44s %8.3f ms  %7.2f T/s  = %6.1f lanes/clk/SM at 1.965 GHz\n", name, ms, ops / ms * 1e-9, ops / (ms * 1e-3) / sms / 1.965e9);
    cudaMemcpy(host, out, (size_t)blocks * 256 * 4, cudaMemcpyDeviceToHost);
    return ms;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, nthr = sms * 4 * 256;
    const uint32_t q = 1073479681u;
    Tw h[8];
    for (int i = 0; i < 8; ++i) {
        const uint64_t w = (123456789ull * (i + 3) * 7919ull) % q;
        const uint64_t kq = (uint64_t)(((unsigned __int128)w << 52) / q);
        h[i].w = (uint32_t)w; h[i].ws = (uint32_t)((w << 32) / q);
        h[i].winv = (double)kq / 4503599627370496.0;
        h[i].c = 6755399441055743.0 - (double)kq;
    }
    Tw *tw; uint32_t *out;
    cudaMalloc(&tw, sizeof(h)); cudaMemcpy(tw, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaMalloc(&out, (size_t)nthr * 4);
    uint32_t *r0 = new uint32_t[nthr], *r1 = new uint32_t[nthr], *r2 = new uint32_t[nthr];
    printf("%s, %d SMs\n", p.name, sms);
    run<3>("IMAD.HI.U32 alone", out, sms, q, tw, r0);
    run<4>("IMAD (32-bit) alone", out, sms, q, tw, r0);
    run<0>("butterfly, quotient by IMAD.HI", out, sms, q, tw, r0);
    run<1>("butterfly, quotient by DADD + DFMA", out, sms, q, tw, r1);
    run<2>("butterfly, quotient by one DFMA", out, sms, q, tw, r2);
    long bad1 = 0, bad2 = 0;
    for (int i = 0; i < nthr; ++i) { bad1 += r0[i] != r1[i]; bad2 += r0[i] != r2[i]; }
    printf("mismatches against the IMAD.HI version (residues compared fully reduced): %ld, %ld of %d threads\n", bad1, bad2, nthr);
    return (bad1 || bad2) ? 1 : 0;
}
