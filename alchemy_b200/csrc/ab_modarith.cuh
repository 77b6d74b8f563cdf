// ab_modarith.cuh -- 32-bit modular arithmetic on the device.  All moduli are primes q < 2^31
// (SURVEY.md F9), residues travel as int64 in HBM (Lol `ZqBasic q Int64`) and as uint32 on chip.
#pragma once
#include <stdint.h>

// Shoup multiplication by a constant w with companion w' = floor(w * 2^32 / q):
// result in [0, 2q) for ANY 32-bit x (needs 2q < 2^32).  3 IMAD-class instructions.
__device__ __forceinline__ uint32_t ab_shoup_lazy(uint32_t x, uint32_t w, uint32_t wp, uint32_t q) {
    return x * w - __umulhi(x, wp) * q;
}
// unsigned "subtract q if >= q": r - q wraps to a huge value when r < q, so min() picks r.
__device__ __forceinline__ uint32_t ab_csub(uint32_t r, uint32_t q) { return min(r, r - q); }
__device__ __forceinline__ uint32_t ab_shoup(uint32_t x, uint32_t w, uint32_t wp, uint32_t q) {
    return ab_csub(ab_shoup_lazy(x, w, wp, q), q);
}
__device__ __forceinline__ uint32_t ab_shoup(uint32_t x, uint2 w, uint32_t q) { return ab_shoup(x, w.x, w.y, q); }
__device__ __forceinline__ uint32_t ab_addmod(uint32_t a, uint32_t b, uint32_t q) { return ab_csub(a + b, q); }
// a - b mod q for a, b in [0,q): a - b wraps when a < b; then (a - b) + q is the small one.
__device__ __forceinline__ uint32_t ab_submod(uint32_t a, uint32_t b, uint32_t q) {
    uint32_t d = a - b;
    return min(d, d + q);
}
// general product of two residues: Barrett on the 62-bit product with mu = floor(2^64 / q)
__device__ __forceinline__ uint32_t ab_mulmod(uint32_t a, uint32_t b, uint32_t q, uint64_t mu) {
    uint64_t x = (uint64_t)a * b;
    uint64_t t = __umul64hi(x, mu);
    uint32_t r = (uint32_t)x - (uint32_t)t * q;   // x - t*q in [0, 2q): the low 32 bits are exact
    return ab_csub(r, q);
}
// centred lift of r in [0,q) to (-q/2, q/2): Lol `lift` (r if 2r < q else r - q)
__device__ __forceinline__ int32_t ab_center(uint32_t r, uint32_t q) {
    return (2u * r < q) ? (int32_t)r : (int32_t)(r - q);   // q < 2^31 so 2r fits in 32 bits
}
// reduce a signed value |v| < 2^31 into [0, q')
__device__ __forceinline__ uint32_t ab_reduce_signed(int32_t v, uint32_t q) {
    int32_t r = v % (int32_t)q;
    return (uint32_t)(r < 0 ? r + (int32_t)q : r);
}
