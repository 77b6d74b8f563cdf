// ab_wire.cu -- a wire / on-disk format for batches of ring elements, ciphertexts and hints.
// The reference keeps keys and hints as in-memory Haskell `Dynamic`s (Crypto/Alchemy/Interpreter/KeysHints.hs:36-39) and
// has no serialisation; sharding ciphertexts over processes or boxes needs one.  Format: a 136-byte little-endian
// header (include/alchemy_b200.h `ab_wire_header`) followed by one uint32 per residue in the device layout
// [batch][npoly][nlimbs][n] -- residues are below 2^31, so the payload is half the size of the int64
// tensors it replaces on NVLink / PCIe / disk.  The header carries the sum of all residues mod 2^64; unpacking
// recomputes it and refuses truncated or corrupted input.
#include <algorithm>
#include <cstddef>
#include <cstring>

#include "ab_internal.h"

static const uint32_t kWireMagic = 0x57324241u;      // "AB2W"
static const uint16_t kWireVersion = 1;

static int64_t wire_count(const ab_wire_header *h) { return h->batch * h->npoly * h->nlimbs * h->n; }
static bool wire_sane(const ab_wire_header *h) {
    return h->magic == kWireMagic && h->version == kWireVersion && h->kind >= 1 && h->kind <= 4 && h->m >= 1 && h->n >= 1 && h->batch >= 0 &&
           h->npoly >= 1 && h->nlimbs >= 1 && h->nlimbs <= AB_MAX_LIMBS && h->n < (1ll << 24) && h->batch < (1ll << 40) && h->npoly < (1 << 20);
}

extern "C" int64_t ab_wire_bytes(const ab_wire_header *h) {
    if (!h || h->batch < 0 || h->npoly < 1 || h->nlimbs < 1 || h->n < 1) return -1;
    return (int64_t)sizeof(ab_wire_header) + 4 * wire_count(h);
}

__global__ void __launch_bounds__(256) k_wire_pack(uint32_t *__restrict__ out, const int64_t *__restrict__ in, int64_t count, unsigned long long *sum) {
    unsigned long long s = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t v = (uint32_t)in[i];
        out[i] = v;
        s += v;
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(sum, s);
}
__global__ void __launch_bounds__(256) k_wire_unpack(int64_t *__restrict__ out, const uint32_t *__restrict__ in, int64_t count, unsigned long long *sum) {
    unsigned long long s = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t v = in[i];
        out[i] = (int64_t)v;
        s += v;
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(sum, s);
}
static unsigned wire_grid(int64_t count) { return (unsigned)std::max<int64_t>(1, std::min<int64_t>((count + 255) / 256, 148 * 16)); }

extern "C" int ab_wire_pack_dev(void *wire_dev, const ab_wire_header *h, const int64_t *data_dev, void *stream) {
    return ab_guard([&] {
        AB_REQUIRE(wire_dev && h && (data_dev || wire_count(h) == 0), AB_ERR_ARG, "null argument");
        ab_wire_header hd = *h;
        hd.magic = kWireMagic; hd.version = kWireVersion; hd.checksum = 0;
        AB_REQUIRE(wire_sane(&hd), AB_ERR_ARG, "bad wire header");
        AB_REQUIRE(((uintptr_t)wire_dev & 7) == 0, AB_ERR_ARG, "wire buffer must be 8-byte aligned");
        cudaStream_t st = (cudaStream_t)stream;
        // the header is staged through a small device-side copy made by value: cudaMemcpyAsync from pageable memory copies before returning
        AB_CHECK_CUDA(cudaMemcpyAsync(wire_dev, &hd, sizeof(hd), cudaMemcpyHostToDevice, st));
        const int64_t count = wire_count(&hd);
        if (count == 0) return;
        auto *sum = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(wire_dev) + offsetof(ab_wire_header, checksum));
        k_wire_pack<<<wire_grid(count), 256, 0, st>>>(reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(wire_dev) + sizeof(ab_wire_header)), data_dev, count, sum);
        ab_count_launch();
        ab_check_launch("wire_pack");
    });
}

extern "C" int ab_wire_read_header(ab_wire_header *h_out, const void *wire_dev, int64_t wire_bytes) {
    return ab_guard([&] {
        AB_REQUIRE(h_out && wire_dev, AB_ERR_ARG, "null argument");
        AB_REQUIRE(wire_bytes >= (int64_t)sizeof(ab_wire_header), AB_ERR_ARG, "wire buffer shorter than a header");
        AB_CHECK_CUDA(cudaMemcpy(h_out, wire_dev, sizeof(ab_wire_header), cudaMemcpyDeviceToHost));
        AB_REQUIRE(wire_sane(h_out), AB_ERR_ARG, "not an alchemy_b200 wire buffer (magic / version / sizes)");
        AB_REQUIRE(ab_wire_bytes(h_out) == wire_bytes, AB_ERR_ARG, "wire buffer truncated or oversized for its header");
    });
}

extern "C" int ab_wire_unpack_dev(int64_t *data_dev, const void *wire_dev, int64_t wire_bytes, void *stream) {
    return ab_guard([&] {
        ab_wire_header h;
        int rc = ab_wire_read_header(&h, wire_dev, wire_bytes);
        if (rc != AB_OK) throw ab_error{rc, ab_last_error()};
        const int64_t count = wire_count(&h);
        AB_REQUIRE(data_dev || count == 0, AB_ERR_ARG, "null argument");
        if (count == 0) return;
        cudaStream_t st = (cudaStream_t)stream;
        unsigned long long *sum = nullptr;
        AB_CHECK_CUDA(cudaMallocAsync((void **)&sum, 8, st));
        AB_CHECK_CUDA(cudaMemsetAsync(sum, 0, 8, st));
        k_wire_unpack<<<wire_grid(count), 256, 0, st>>>(data_dev, reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(wire_dev) + sizeof(ab_wire_header)), count, sum);
        ab_count_launch();
        ab_check_launch("wire_unpack");
        unsigned long long got = 0;
        AB_CHECK_CUDA(cudaMemcpyAsync(&got, sum, 8, cudaMemcpyDeviceToHost, st));
        AB_CHECK_CUDA(cudaStreamSynchronize(st));
        cudaFreeAsync(sum, st);
        AB_REQUIRE(got == h.checksum, AB_ERR_ARG, "wire checksum mismatch: payload corrupted");
    });
}
