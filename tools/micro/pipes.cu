// Micro-benchmark: per-SM throughput of IMAD.WIDE.U32, DFMA and both at once (half the warps each), sm_100a.
// Question it answers: can exact-integer dot products move from the FMA-heavy pipe (IMAD.WIDE) to the FP64 pipe (DFMA) and run concurrently?
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu ; run: ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>   // 0: IMAD.WIDE, 1: DFMA, 2: warps alternate (warp/4 parity), 3: IMAD (32-bit mad.lo)
__global__ void __launch_bounds__(256) k(uint64_t *out, int iters, uint32_t a, double b) {
    const int w = threadIdx.x >> 5;
    const bool f64 = MODE == 1 || (MODE == 2 && ((w >> 2) & 1));
    uint64_t acc[8];
    double dacc[8];
    uint32_t x = threadIdx.x * 2654435761u + 12345u, xs[8];
    double dx[8];
    for (int i = 0; i < 8; ++i) { acc[i] = i + x; dacc[i] = i + (double)(x & 1023); xs[i] = x + i * 977u; dx[i] = (double)(xs[i] & 0xffff); }
    if (MODE == 3) {
        uint32_t s[8];
        for (int i = 0; i < 8; ++i) s[i] = i + x;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) s[i] = s[(i + 1 + u) & 7] * a + s[i];
        }
        uint64_t r = 0;
        for (int i = 0; i < 8; ++i) r += s[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = r;
        return;
    }
    if (!f64) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] += (uint64_t)(uint32_t)acc[(i + 1 + u) & 7] * a;
        }
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) dacc[i] = fma(dacc[(i + 1 + u) & 7], b, dacc[i]);
        }
    }
    uint64_t r = 0;
    for (int i = 0; i < 8; ++i) r += acc[i] + (uint64_t)dacc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char *name, uint64_t *out, int sms) {
    const int iters = 4096, blocks = sms * 4;
    k<MODE><<<blocks, 256>>>(out, 16, 3, 1.5);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, iters, 3, 1.5);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)blocks * 256 * iters * 32;         // thread-level multiply-adds
    printf("%-28s %8.3f ms  %8.2f T mad/s  = %6.1f lanes/clk/SM at 1.965 GHz\n", name, ms, ops / ms * 1e-9, ops / (ms * 1e-3) / sms / 1.965e9);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    uint64_t *out;
    cudaMalloc(&out, (size_t)p.multiProcessorCount * 4 * 256 * 8);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    run<0>("IMAD.WIDE.U32 (all warps)", out, p.multiProcessorCount);
    run<1>("DFMA (all warps)", out, p.multiProcessorCount);
    run<2>("half IMAD.WIDE, half DFMA", out, p.multiProcessorCount);
    run<3>("IMAD 32-bit (all warps)", out, p.multiProcessorCount);
    run<0>("IMAD.WIDE.U32 again", out, p.multiProcessorCount);
    return 0;
}
